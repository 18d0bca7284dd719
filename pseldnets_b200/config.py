"""Factory seam: the one place the reference wires its extractor
(/root/reference/src/utils/config.py:24-32, `get_afextractor(cfg)` keyed on
`cfg['data']['audio_feature']`).  Same name, same argument, same return convention (an nn.Module, or
None for feature kinds that have no on-line extractor)."""
from . import feature

# audio_feature value -> extractor class.  'logmelgcc' is an addition: the reference returns None for it and
# computes MIC features offline (src/preproc/preprocess.py:525-563).
_EXTRACTORS = {
    'logmelIV': feature.LogmelIV_Extractor,
    'logmel': feature.Logmel_Extractor,
    'logmelgcc': feature.LogmelGCC_Extractor,
}


def get_afextractor(cfg):
    """Audio feature extractor for this config, or None (e.g. 'salsalite')."""
    cls = _EXTRACTORS.get(cfg['data']['audio_feature'])
    return cls(cfg) if cls is not None else None
