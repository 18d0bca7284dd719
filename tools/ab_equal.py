import os, sys, subprocess, hashlib
code = r'''
import sys, torch, hashlib
sys.path.insert(0, ".")
import pseldnets_b200 as pb
def cfg(feat, sr=24000, hop=240): return {"data": {"sample_rate": sr, "nfft": 1024, "hoplen": hop, "n_mels": 64, "window": "hann", "audio_feature": feat}}
torch.manual_seed(0)
h = hashlib.sha256()
x = 0.1 * torch.randn(3, 4, 48000, device="cuda")
for feat, xx in (("logmelIV", x), ("logmelIV", torch.randn(2, 8, 20000, device="cuda")), ("logmel", x[:, :3]), ("logmel", x[:, :1])):
    y = pb.get_afextractor(cfg(feat)).cuda()(xx)
    h.update(y.cpu().numpy().tobytes())
y = pb.get_afextractor(cfg("logmelIV", 32000, 320)).cuda()(x); h.update(y.cpu().numpy().tobytes())
print(h.hexdigest())
'''
for v in "AB":
    env = dict(os.environ, SELD_LIB=os.path.abspath("build/ab/lib%s.so" % v))
    print(v, subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip())
