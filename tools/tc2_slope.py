"""Main-loop rate of the pair engine (engine 2) vs the number of active CTA pairs, operands L2-RESIDENT (A + B images
well under 126 MB) so that the slope over K isolates the L2 -> SM feed + MMA issue from HBM streaming; and one
HBM-streaming case for contrast.  Intercept = per-launch cost outside the main loop."""
import sys
sys.argv = ['x']
exec(open('tools/tc_microbench.py').read().split("KK = (True, True)")[0])
cases = [(2048, 1024, 256, 1024, "32 pairs, 1 wave"), (4096, 1024, 256, 1024, "64 pairs, 1 wave"), (9472, 512, 256, 1024, "74 pairs, 1 wave"),
         (9472, 1024, 256, 1024, "148 pairs, 2 waves"), (7500, 1000, 392, 784, "cfg3 R-op 1 shape: 120 pairs, 2 waves"),
         (9472, 2048, 1024, 4096, "296 pairs, 4 waves, operands 380 MB: streams from HBM")]
for (M, N, K0, K1, what) in cases:
    for eng in (2, 1):
        t0, t1 = bench(M, N, K0, eng), bench(M, N, K1, eng)
        units = 74 if eng == 2 else 148
        tiles = (-(-M // 256)) * (-(-N // 256)) if eng == 2 else (-(-M // 128)) * (-(-N // 128))
        waves = -(-tiles // units)
        slope = (t1 - t0) / ((K1 - K0) / 32) / waves
        print(f"{what:58s} engine {eng}: K={K0} {t0:6.1f} us, K={K1} {t1:6.1f} us -> {slope:.3f} us per k-block and wave, "
              f"intercept {t0 - slope * waves * K0 / 32:5.1f} us", flush=True)
