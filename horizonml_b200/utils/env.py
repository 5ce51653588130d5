"""Environment report: ``python -m horizonml_b200.utils.env`` (what is built, what is visible)."""
from __future__ import annotations

import json
import os


def report() -> dict:
    import torch
    from ..ops import _ext
    info = {"torch": torch.__version__, "cuda_available": torch.cuda.is_available(),
            "gpus": torch.cuda.device_count() if torch.cuda.is_available() else 0,
            "extension_built": _ext.is_built(), "extension_path": _ext._SO if os.path.exists(_ext._SO) else None,
            "nvcc": _ext._nvcc(), "source_hash": _ext.source_hash(),
            "env": {k: os.environ[k] for k in ("HZ_PDL", "HZ_SPLITK", "HZ_CLUSTER_SPLITK", "HZ_DISABLE_NVLS",
                                               "HZ_STRICT_NATIVE") if k in os.environ}}
    if info["cuda_available"]:
        p = torch.cuda.get_device_properties(0)
        info["gpu0"] = {"name": p.name, "sm": f"{p.major}.{p.minor}", "sms": p.multi_processor_count,
                        "mem_gb": round(p.total_memory / 2**30, 1)}
    return info


if __name__ == "__main__":
    print(json.dumps(report(), indent=1))
