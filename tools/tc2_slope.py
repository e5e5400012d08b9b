import sys; sys.argv=['x']
exec(open('tools/tc_microbench.py').read().split("KK = (True, True)")[0])
for (M,N,what) in [(2048,1024,"32 pairs"),(4096,1024,"64 pairs"),(4736,1024,"74 pairs = 1 wave"),(9472,2048,"296 pairs = 4 waves")]:
    ts=[bench(M,N,K,2) for K in (1024,4096)]
    print(f"{what}: K=1024 {ts[0]:.1f} us, K=4096 {ts[1]:.1f} us, slope {(ts[1]-ts[0])/96/max(1,round(M/256*N/256/74)):.3f} us per k-block")
