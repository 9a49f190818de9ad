import sys; sys.path.insert(0,"othello-alphazero_b200"); sys.path.insert(0,"oracle"); sys.path.insert(0,"tests")
import numpy as np, othello_mcts as om, pyoracle as po
from helpers import playout_positions
L=po.lib()
rows=playout_positions(po, 1500, seed=2024)
pos=[r[0] for r in rows]; acts=[r[2] for r in rows]
me=np.array([p.p1 if p.player==1 else p.p2 for p in pos],np.uint64); opp=np.array([p.p2 if p.player==1 else p.p1 for p in pos],np.uint64)
moves=np.array([(1<<(63-a)) if a!=64 else 0 for a in acts],np.uint64)
for n in (3018, 20000, len(moves)):
    fl=om.flips_batch(moves[:n],me[:n],opp[:n])
    want=np.array([L.ozo_get_flips(int(m),int(a),int(b)) for m,a,b in zip(moves[:n],me[:n],opp[:n])],np.uint64)
    bad=np.flatnonzero(fl!=want); print("n",n,"bad",len(bad), bad[:10])
    for i in bad[:6]:
        print(i, acts[i], hex(int(moves[i])), hex(int(me[i])), hex(int(opp[i])), "gpu", hex(int(fl[i])), "want", hex(int(want[i])), "host", hex(om.get_flips(int(moves[i]),int(me[i]),int(opp[i]))))
