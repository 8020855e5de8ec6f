import os, sys, torch
sys.path.insert(0, "/root/repo")
from echo_tts_b200 import _lib
_lib.load(strict=False)
from echo_tts_b200.autoencoder import B200DAC, PCAState, ae_decode
from echo_tts_b200.config import DacConfig
from echo_tts_b200.weights import make_dac_weights, make_pca_state
g = torch.load("/root/repo/tests/golden/dac_full_T64.pt", weights_only=True)
cfg = DacConfig.base()
dac = B200DAC.from_state_dict(make_dac_weights(cfg, seed=4321), cfg, "cuda:0")
comps, mean, scale = make_pca_state(cfg)
a = ae_decode(dac, PCAState(comps, mean, scale), g["z"]).cpu()
print(os.environ.get("ECHO_B200_LIB", "current"), "T=64 rel-L2", ((a - g["audio"]).norm() / g["audio"].norm()).item())
