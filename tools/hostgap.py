import time, torch, numpy as np, sys
sys.path.insert(0, '.')
import thesia_b200 as thb
_st = torch.cuda.Stream(); torch.cuda.set_stream(_st)
ctx = thb.Context(0, _st.cuda_stream)
sr, n, nch = 48000, 48000*600, 128
pcm = torch.empty((nch, n), dtype=torch.float32, device='cuda')
for c in range(nch): ctx.synth_pcm(pcm[c], sr, c//2, c%2, 0)
ctx.synchronize()
s = thb.SpecSetting(2048/48.0, 4, 1, thb.FreqScale.Mel, 128)
tracks = ctx.prepare_tracks([dict(pcm=pcm[c], id=c//2, ch=c%2, sr=sr) for c in range(nch)])
for _ in range(3):
    ctx.spec_batch(tracks, s); ctx.update_spec_imgs(100.0, 258, sr)
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); ctx.spec_batch(tracks, s); t1 = time.perf_counter()
    r = ctx.update_spec_imgs(100.0, 258, sr); t2 = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"spec_batch host {1e3*(t1-t0):.3f} ms, update_spec_imgs (incl. wait) {1e3*(t2-t1):.3f} ms, tail {1e3*(t3-t2):.3f}")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): ctx.spec_batch(tracks, s)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(8)
