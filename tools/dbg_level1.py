import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import maelstrom_b200 as mb
from maelstrom_b200.engine import KIND_SIM_CLIENT
import oracle_lib as O
from scenarios import random_broadcast_ops
n=1024; V=int(sys.argv[1]) if len(sys.argv)>1 else 6000
g = mb.Sim(n, workload="broadcast", topology="grid", n_values=V+8, ring_cap=8192, max_window=4096, journal_cap_log2=25, journal_level=1, max_endpoints=n+32)
o = O.Sim(n, workload=O.W_BROADCAST, topology="grid", n_values=V+8)
cg=[g.add_endpoint("c%d"%i, KIND_SIM_CLIENT) for i in range(16)]
co=[o.add_endpoint("c%d"%i, O.KIND_SIM_CLIENT) for i in range(16)]
ops,nv = random_broadcast_ops(n, cg, n_ticks=1, per_tick=V, seed=5)
g.schedule(ops); o.schedule(ops)
g.phase_cycles(True)
t=time.time(); g.run(1_000_000); print("gpu", time.time()-t)
pc=g.phase_cycles(True)
for c in range(4):
    print(c, int(pc[c][15]), [int(pc[c][k]) for k in (9,10,11,12,13)])
t=time.time(); o.run(1_000_000); print("oracle", time.time()-t)
ev_g,_ = g.drain(bodies=False)
ev_o,_ = o.journal()
print(len(ev_g), len(ev_o), g.stats()==o.stats(), g.counters())
ok=True
for f in ("event_id","time_ns","msg_id","src","dest"):
    same = np.array_equal(ev_g[f], ev_o[f]); ok &= same
    if not same:
        bad=int(np.nonzero(ev_g[f]!=ev_o[f])[0][0]); print(f, "differs at", bad, ev_g[bad], ev_o[bad])
print("PARITY", ok)
