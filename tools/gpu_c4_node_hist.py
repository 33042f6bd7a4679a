import sys, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import import_package
import_package()
from nori_ray_tracer_b200 import host_scene, abi
from nori_ray_tracer_b200.gpu import NoriGpu
sc = host_scene.heightfield_scene(n=2237)
g = NoriGpu(0); g.upload_scene(sc)
cam = sc.pod.camera
s2c = np.array(cam.sampleToCamera[:], np.float64).reshape(4, 4); c2w = np.array(cam.cameraToWorld[:], np.float64).reshape(4, 4)
rng = np.random.RandomState(0); N = 400000
px = rng.rand(N) ; py = rng.rand(N)
p = np.stack([px, py, np.zeros(N), np.ones(N)], 0); q = s2c @ p; near = (q[:3] / q[3]).T
d = near / np.linalg.norm(near, axis=1, keepdims=True); dw = (c2w[:3, :3] @ d.T).T; o = c2w[:3, 3]
rays = np.zeros(N, abi.RAY_DTYPE); rays['o'] = o; rays['d'] = dw; rays['mint'] = 1e-4 / d[:, 2]; rays['maxt'] = 1e4 / d[:, 2]
h = g.trace(rays, 0)
nv = h['nodes_visited'].astype(np.int64)
print('primary: mean', nv.mean(), 'pcts', np.percentile(nv, [50, 90, 99, 99.9, 100]), 'hit frac', (h['shape'] != 0xffffffff).mean(), 'trace ms', g.stats().trace_ms)
# secondary rays from hit points, random hemisphere directions
hit = h['shape'] != 0xffffffff
P = rays['o'][hit] + rays['d'][hit] * h['t'][hit][:, None]
M = len(P); dd = rng.randn(M, 3); dd /= np.linalg.norm(dd, axis=1, keepdims=True); dd[:, 2] = np.abs(dd[:, 2])
r2 = np.zeros(M, abi.RAY_DTYPE); r2['o'] = P; r2['d'] = dd; r2['mint'] = 1e-4; r2['maxt'] = np.inf
h2 = g.trace(r2, 0); nv2 = h2['nodes_visited'].astype(np.int64)
print('secondary: mean', nv2.mean(), 'pcts', np.percentile(nv2, [50, 90, 99, 99.9, 100]), 'trace ms', g.stats().trace_ms)
# per-warp max/mean
for name, v in (('primary', nv), ('secondary', nv2)):
    w = v[: len(v) // 32 * 32].reshape(-1, 32)
    print(name, 'mean of warp-max / mean', w.max(1).mean() / v.mean())
