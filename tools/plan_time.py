import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import b2gpkg
b2g=b2gpkg.load()
ctx=b2g.Context(0)
for path in ['tests/golden/h10_sz_m40_s4.b2seq','workloads/cr2_svp_m4000_site20.b2seq.gz']:
    sf=b2g.load_seqfile(path)
    if sf.arenas is None:
        sf=sf.subset(np.arange(sf.npairs)%4==0)   # quarter of the Cr2 list: ~3 GB of operands from host memory
        sf.arenas=np.random.default_rng(0).standard_normal(sf.operand_doubles)
    for rep in range(3):
        t=time.perf_counter(); plan=b2g.SeqPlan.from_seqfile(ctx,sf,sf.arenas); dt=time.perf_counter()-t
        print(path, 'pairs',sf.npairs,'operand GB %.2f'%(sf.operand_doubles*8e-9),'plan_create %.1f ms'%(dt*1e3), 'upload_s', plan.stats.upload_seconds, flush=True)
        plan.close()
