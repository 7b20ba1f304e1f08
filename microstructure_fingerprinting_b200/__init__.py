"""B200-native (sm_100a) microstructure fingerprinting: the per-voxel exhaustive dictionary
fit of rensonnetg/microstructure_fingerprinting behind the reference's own Python API.

    from microstructure_fingerprinting_b200 import MFModel
    fit = MFModel(dictionary).fit(data, mask, numfasc, peaks=..., bvals=..., bvecs=...)
    fit.write_nifti('out/subject')

The hot path (dictionary rotation, Gram contractions, combinatorial search) runs in
hand-written CUDA kernels (libmfb200.so, C ABI in include/mfb200.h); there is no CPU
fallback.
"""
from .mf import MFModel, MFModelFit
from .peaks import cleanup_2fascicles
from . import mf_utils

__version__ = "0.1.0"
__all__ = ["MFModel", "MFModelFit", "cleanup_2fascicles", "mf_utils"]
