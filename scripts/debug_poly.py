"""Debug helper (GPU box): compare GPU and oracle output per polygon and print the differing ones."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import myrenderer_b200 as mr
from oracle import oracle as O

seed = 0x5EED0003
npoly = 4000
fp = O.synth_polygon_sizes(seed, npoly, 8, 64)
xy = O.synth_polygons(seed, fp)
ctx = mr.Context(0)
b = mr.Polygon(ctx).create_polygons(xy, fp, seed=seed)
ctx.sync()
ref = O.polygon_batch(xy, fp, seed=seed)
gs = b.status.cpu().numpy().view(np.uint32)
gv = b.vertex_buffer.cpu().numpy()
ft = ref["first_tri"]
bad = np.where(gs != ref["status"])[0]
print("bad", bad)
for i in bad[:6]:
    n = int(fp[i + 1] - fp[i])
    P = xy[fp[i]:fp[i + 1]]
    r1 = O.polygon_batch(P, np.array([0, n]), seed=seed, poly_index0=int(i), want_stats=True)
    st = r1["stats"]
    print(f"poly {i} n={n} gpu_status={gs[i]} oracle={ref['status'][i]} nodes={st['nodes']} (tier0 cap {6*(16<<[c for c in range(8) if n<=16<<c][0])+32})"
          f" max_stack={st['max_stack']} mountains={st['mountains']} tris={st['triangles']}")
    a, z = int(ft[i]) * 96, int(ft[i + 1]) * 96
    g = gv[a:z].reshape(-1, 32)[:, :8].copy().view(np.float32).reshape(-1, 2)
    ids = []
    for q in g:
        m = np.where((P[:, 0] == q[0]) & (P[:, 1] == q[1]))[0]
        ids.append(int(m[0]) if len(m) else -1)
    print(" gpu ids   ", np.array(ids).reshape(-1, 3).tolist())
    print(" oracle ids", ref["ids"][3 * int(ft[i]):3 * int(ft[i + 1])].reshape(-1, 3).astype(np.int64).tolist())
    # alone
    b1 = mr.Polygon(ctx).create_polygons(P, np.array([0, n], dtype=np.uint64), seed=seed, poly_index0=int(i))
    ctx.sync()
    print(" alone status", b1.status.cpu().numpy())
