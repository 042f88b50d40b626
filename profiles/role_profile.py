import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from audiality2_b200 import engine as eng
from audiality2_b200.workloads import cfg2_bank
from audiality2_b200.chains import autowire
import numpy as np
e=eng.Engine(48000,2); b=cfg2_bank(4096); w=e.builtin_wave('saw')
bank=e.new_bank(autowire(list(b['kinds'])),4096)
e.write_all(bank,0,0,[w<<16]); e.write_all(bank,0,1,b['pitch']); e.write_all(bank,0,2,[b['amp']])
e.write_all(bank,1,0,b['cutoff']); e.write_all(bank,1,1,[b['q']]); e.write_all(bank,2,1,b['pan'])
e.set_timing(True)
for i in range(5): e.write_all(bank,0,2,[b['amp']//(1+i%2)],dur=960<<8); e.run(960,64)
e.split_profile(True, False)
N=20
for i in range(N): e.write_all(bank,0,2,[b['amp']//(1+i%2)],dur=960<<8); e.run(960,64)
p=e.split_profile(False, True)
ncta=128
print('kernel ms', e.last_render_ms(), 'split launches', e.split_launches)
names=['control','serial','stageA','stageC','barrier-wait(helper)','iters']
it=p[5]
for n,v in zip(names,p): print('%-22s %10.0f cycles per CTA-iteration'%(n, v/max(it,1)))
nl = e.split_launches - 5
print('per CTA and launch: prologue %.0f cycles, pipeline + state store %.0f cycles (%d launches x %d CTAs)' % (p[6]/(N*ncta), p[7]/(N*ncta), N, ncta))
