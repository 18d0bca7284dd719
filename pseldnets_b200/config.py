"""Factory seam of the reference: /root/reference/src/utils/config.py:24-32."""
from . import feature


def get_afextractor(cfg):
    """ Get audio feature extractor."""
    if cfg['data']['audio_feature'] == 'logmelIV':
        afextractor = feature.LogmelIV_Extractor(cfg)
    elif cfg['data']['audio_feature'] == 'logmel':
        afextractor = feature.Logmel_Extractor(cfg)
    else:
        afextractor = None
    return afextractor
