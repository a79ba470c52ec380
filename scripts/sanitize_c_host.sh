#!/bin/bash
# compute-sanitizer over the plain-C host (examples/c_host.c): no Python or torch in the process, so every report is
# about libgsvc_rast.so — initcheck in particular (torch's caching allocator hides uninitialised reads from it).
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
CUDA=${CUDA_HOME:-/usr/local/cuda}
TMP=$(mktemp -d)
gcc -std=c99 -O1 -I"$ROOT/include" -I"$CUDA/include" "$ROOT/examples/c_host.c" -o "$TMP/c_host" \
    -L"$ROOT/gsvc_b200" -lgsvc_rast -L"$CUDA/lib64" -lcudart -Wl,-rpath,"$ROOT/gsvc_b200" -Wl,-rpath,"$CUDA/lib64" || exit 1
cd "$ROOT" && python - "$TMP/scene.bin" <<'PY'
import sys
import numpy as np
from tests.scenes import make_scene, np_inputs
P, W, H = 6000, 150, 90
scene = make_scene(P=P, W=W, H=H, F=128, seed=31, back=True, bg=(0.3, 0.1, 0.6), scale_modifier=0.5)
st, gi = scene["oracle_settings"], np_inputs(scene["gaussians"])
with open(sys.argv[1], "wb") as f:
    np.asarray([W, H, P], np.int32).tofile(f)
    np.asarray([st.x_min, st.y_min, st.scale, st.threshold, st.scale_modifier], np.float32).tofile(f)
    np.asarray(st.bg, np.float32).tofile(f)
    np.asarray(st.viewmatrix, np.float32).reshape(16).tofile(f)
    for k in ("means3D", "opacities", "colors_precomp", "scales", "rotations"):
        np.ascontiguousarray(gi[k], np.float32).tofile(f)
    np.random.default_rng(5).standard_normal((3, H, W)).astype(np.float32).tofile(f)
PY
for tool in memcheck initcheck synccheck racecheck; do
  echo "== $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 "$TMP/c_host" "$TMP/scene.bin" "$TMP/out_$tool.bin" 2>&1 \
    | grep -E "SUMMARY|num_rendered|Uninitialized|Invalid|Error|hazard" | head -12
done
# num_rendered + image + radii are deterministic (the gradients behind them depend on the order of the float atomics)
cmp -n $((8 + 12 * 150 * 90 + 4 * 6000)) "$TMP/out_memcheck.bin" "$TMP/out_initcheck.bin" && echo "count, image and radii identical across runs"
rm -rf "$TMP"
