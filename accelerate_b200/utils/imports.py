"""Availability probes user scripts written for the reference call (``src/accelerate/utils/imports.py``).  This
build targets one platform — CUDA on B200 plus the CPU/gloo debugging path — so every other accelerator answers no."""

import importlib.util

import torch

from .environment import patch_environment


def _has(module: str) -> bool:
    return importlib.util.find_spec(module) is not None


def is_cuda_available() -> bool:
    """NVML-based check that does not initialise the driver (imports.py:111-119)."""
    with patch_environment(PYTORCH_NVML_BASED_CUDA_CHECK="1"):
        return torch.cuda.is_available()


def is_bf16_available(ignore_tpu: bool = False) -> bool:
    return torch.cuda.is_bf16_supported() if is_cuda_available() else True


def is_fp16_available() -> bool:
    return True


def is_datasets_available() -> bool:
    return _has("datasets")


def is_pytest_available() -> bool:
    return _has("pytest")


def is_safetensors_available() -> bool:
    return _has("safetensors")


def is_transformers_available() -> bool:
    return _has("transformers")


def _no(*args, **kwargs) -> bool:
    return False


# accelerators and backends this build does not drive
is_mps_available = is_hpu_available = is_xpu_available = is_npu_available = is_mlu_available = is_musa_available = _no
is_sdaa_available = is_neuron_available = is_torch_xla_available = is_deepspeed_available = is_megatron_lm_available = _no
is_fp8_available = is_torchao_available = is_transformer_engine_available = is_msamp_available = is_bnb_available = _no


def is_torchdata_stateful_dataloader_available() -> bool:
    """torchdata >= 0.8 ships ``torchdata.stateful_dataloader.StatefulDataLoader``."""
    try:
        return importlib.util.find_spec("torchdata.stateful_dataloader") is not None
    except ModuleNotFoundError:
        return False
