"""Runs in the BUILD container only (needs /root/reference): the reference DAC decode in bf16 -- the precision handler.py:379-384
runs it in -- against the reference in fp32 on the same weights and latents. Context for the audio tolerance: the reference
itself moves by 3.3e-2 rel-L2 when it is run in bf16; this library (bf16 operands, fp32 accumulation and stream) by 1.6e-2."""
import sys, os, time, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import pin_reference as PR
from echo_tts_b200.config import DacConfig
from echo_tts_b200.weights import make_dac_weights, make_pca_state
refmodel, refinf, refblk, refae = PR.import_reference()
torch.set_num_threads(8)
dcfg = DacConfig.base()
dsd = make_dac_weights(dcfg, seed=4321)
ae = PR.build_ref_dac(refae, dcfg, dsd)
comps, mean, scale = make_pca_state(dcfg)
pca = refinf.PCAState(pca_components=comps, pca_mean=mean, latent_scale=scale)
g = torch.load("/root/repo/tests/golden/dac_full_T64.pt", weights_only=True)
z = g["z"][:, :32]
with torch.inference_mode():
    a32 = refinf.ae_decode(ae, pca, z)
    print("fp32 vs golden prefix", ((a32 - g["audio"][..., :32*2048]).norm() / g["audio"][..., :32*2048].norm()).item())
    t0 = time.time()
    ae16 = ae.to(torch.bfloat16)
    pm = ae16.quantizer.post_module
    a16 = refinf.ae_decode(ae16, pca, z)
    print("reference bf16 decode", time.time() - t0, "s")
    print("REFERENCE bf16 (handler.py:379-384 runs fish_ae in bf16) vs REFERENCE fp32, T=32: rel-L2", ((a16.float() - a32).norm() / a32.norm()).item())
