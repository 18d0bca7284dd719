"""B200-native SELD feature front-end: a drop-in for the extractors of Jinbo-Hu/PSELDNets
(`src/utils/feature.py`), computed by fused sm_100a CUDA kernels behind a C ABI."""
from . import augment
from .config import get_afextractor
from .epilogue import ScalarParams, apply_scalar, reshape_wav2img, scalar_wav2img
from .graphs import GraphedFrontEnd
from .feature import Features_Extractor_MIC, LogmelGCC_Extractor, LogmelIV_Extractor, Logmel_Extractor

__all__ = ['LogmelIV_Extractor', 'Logmel_Extractor', 'LogmelGCC_Extractor', 'Features_Extractor_MIC',
           'get_afextractor', 'augment', 'GraphedFrontEnd', 'ScalarParams', 'apply_scalar', 'reshape_wav2img', 'scalar_wav2img']
__version__ = '0.1'
