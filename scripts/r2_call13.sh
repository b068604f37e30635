#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest13.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest13.log
python scripts/r2_probe.py cfg4 kd 2>&1 | tail -9
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import ray_tracing_engine_b200 as rt
g = np.load('tests/golden/photons.npz')
r = rt.Renderer(rt.Scene.load('tests/golden/scenes/stock.rtscene'), 1, 0, None, 3000, 10, seed=1)
pl, counts, hist = r.emit_photons()
w = g['list']
print('stock 3000: gpu', len(pl), 'ref', len(w), 'identical rows', (pl.view(np.uint32) == w.view(np.uint32)).all(1).mean() if pl.shape == w.shape else 'shape differs', 'hist equal', (hist == g['hist']).all())
g2 = np.load('tests/golden/render_example_m1_N128_p50000_k10_win.npz')
r = rt.Renderer(rt.Scene.load('tests/golden/scenes/example.rtscene'), 1, 0, None, 50000, 10, seed=1)
pl, counts, hist = r.emit_photons()
w = g2['photons']
print('example 50000: gpu', len(pl), 'ref', len(w), 'identical rows', (pl.view(np.uint32) == w.view(np.uint32)).all(1).mean() if pl.shape == w.shape else 'shape differs', 'hist equal', (hist == g2['depth_hist']).all())
PY
