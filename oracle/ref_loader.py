"""ORACLE — TEST INFRASTRUCTURE. Imports the UNMODIFIED reference (vilbert/vilbert.py) from
/root/reference in the build container (it does not exist on the GPU box), with the four stub modules
the survey found necessary under torch 2.x: boto3, botocore.exceptions, tensorboardX,
torch._six. Used only by oracle/make_golden.py to pin oracle/vilbert_oracle.py and to write fixtures."""
import math
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VILBERT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "vilbert", "vilbert.py"))


def load():
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "boto3" not in sys.modules:
        sys.modules["boto3"] = types.ModuleType("boto3")
    if "botocore" not in sys.modules:
        bc = types.ModuleType("botocore"); ex = types.ModuleType("botocore.exceptions")
        ex.ClientError = type("ClientError", (Exception,), {})
        bc.exceptions = ex
        sys.modules["botocore"] = bc; sys.modules["botocore.exceptions"] = ex
    if "tensorboardX" not in sys.modules:
        tb = types.ModuleType("tensorboardX"); tb.SummaryWriter = type("SummaryWriter", (), {})
        sys.modules["tensorboardX"] = tb
    if "torch._six" not in sys.modules:
        six = types.ModuleType("torch._six"); six.inf = math.inf
        sys.modules["torch._six"] = six
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import vilbert.vilbert as ref  # noqa: E402
    return ref
