import os, sys, subprocess, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    from echo_tts_b200.autoencoder import B200DAC
    from echo_tts_b200.config import DacConfig
    from echo_tts_b200.weights import make_dac_weights
    cfg = DacConfig.base()
    dac = B200DAC.from_state_dict(make_dac_weights(cfg, 4321), cfg, "cuda:0")
    g = torch.Generator().manual_seed(3)
    zq = torch.randn(2, 1024, 128, generator=g).cuda()
    outs = {}
    for name, z in (("b2", zq), ("b1", zq[:1]), ("b2_again", zq), ("b1_T64", zq[:1, :, :64])):
        outs[name] = dac.decode_zq(z).float().cpu()
    torch.save(outs, sys.argv[1])
else:
    cfgs = (("nosplit", {"ECHO_SPLIT_K": "1"}), ("split", {}), ("split_nopdl", {"ECHO_NO_PDL": "1"}),
            ("split2", {"ECHO_SPLIT_K": "2"}), ("split_blocking", {"CUDA_LAUNCH_BLOCKING": "1"}))
    res = {}
    for tag, env in cfgs:
        subprocess.run([sys.executable, __file__, f"/tmp/dac_{tag}.pt"], env=dict(os.environ, **env), check=True)
        res[tag] = torch.load(f"/tmp/dac_{tag}.pt")
    a = res["nosplit"]
    rel = lambda x, y: ((x - y).norm() / y.norm()).item()
    for tag, o in res.items():
        print(tag, "b2 vs nosplit b2:", "%.3e" % rel(o["b2"], a["b2"]))
        print(tag, "b2[0] vs b1:", "%.3e" % rel(o["b2"][0], o["b1"][0]), " b2 vs b2_again:", "%.3e" % rel(o["b2"], o["b2_again"]),
              " b1[:64] prefix vs b1_T64:", "%.3e" % rel(o["b1"][0, :, :64 * 2048], o["b1_T64"][0]))
