"""Platform check.  The reference lazily detects NVIDIA/AMD/XPU/CPU (conch/platforms/platform.py:70-87);
here there is exactly one platform: CUDA on an sm_100 (B200) device -- anything else raises."""

from __future__ import annotations

import functools

import torch


class _Platform:
    device = "cuda"

    def is_nvidia(self) -> bool:
        return True

    def has_cuda(self) -> bool:
        return True

    def is_amd(self) -> bool:
        return False

    def is_xpu(self) -> bool:
        return False

    def is_unspecified(self) -> bool:
        return False

    def supports_fp8(self) -> bool:
        return True  # sm_100: e4m3 / e5m2 tensor-core operands (reference conch/platforms/platform.py:50-60)

    def name(self) -> str:
        return "nvidia"

    @functools.lru_cache(maxsize=None)  # noqa: B019
    def check(self, index: int | None = None) -> None:
        if not torch.cuda.is_available():
            raise RuntimeError("conch_b200 needs a CUDA device (B200, sm_100); there is no CPU fallback")
        major, minor = torch.cuda.get_device_capability(index)
        if major != 10:
            raise RuntimeError(f"conch_b200 is built for sm_100a only, found sm_{major}{minor}")

    def get_device_name(self) -> str:
        return torch.cuda.get_device_name()


current_platform = _Platform()
