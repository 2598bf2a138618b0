"""admm_b200 -- B200-native (sm_100a) ADMM solvers behind the API of the R package yixuan/ADMM.

csrc/      hand-written CUDA kernels + the C ABI (include/b200admm.h) -> libb200admm.so
_capi.py   ctypes binding of that ABI
api.py     the reference's admm_lasso / admm_enet / admm_lad / admm_bp chain, restated
build.py   nvcc build recipe
"""
from .api import (ADMM_BP, ADMM_BP_fit, ADMM_Enet, ADMM_Enet_fit, ADMM_LAD, ADMM_LAD_fit, ADMM_Lasso,
                  ADMM_Lasso_fit, admm_bp, admm_enet, admm_lad, admm_lasso)
from ._capi import B200AdmmError, device_info

__all__ = ["admm_lasso", "admm_enet", "admm_lad", "admm_bp", "ADMM_Lasso", "ADMM_Enet", "ADMM_LAD", "ADMM_BP",
           "ADMM_Lasso_fit", "ADMM_Enet_fit", "ADMM_LAD_fit", "ADMM_BP_fit", "B200AdmmError", "device_info"]
