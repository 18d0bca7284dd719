"""Factory seam of the reference: /root/reference/src/utils/config.py:24-32."""
from . import feature


def get_afextractor(cfg):
    """ Get audio feature extractor."""
    if cfg['data']['audio_feature'] == 'logmelIV':
        afextractor = feature.LogmelIV_Extractor(cfg)
    elif cfg['data']['audio_feature'] == 'logmel':
        afextractor = feature.Logmel_Extractor(cfg)
    elif cfg['data']['audio_feature'] == 'logmelgcc':
        # the reference returns None here (MIC features come from its offline numpy class);
        # this build adds the on-GPU module for the same features
        afextractor = feature.LogmelGCC_Extractor(cfg)
    else:
        afextractor = None
    return afextractor
