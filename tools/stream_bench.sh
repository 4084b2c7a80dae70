#!/bin/bash
# f-2 data point: text-protocol throughput of interactive_mode, reference CLI (CPU) vs the streaming tool (GPU),
# on the reference-trained multi-simple snapshot (n=100, nt=6, nr=5).
N=${1:-200000}
python - <<PY
import numpy as np
rng = np.random.default_rng(1)
np.savetxt("/tmp/stream_pts.txt", rng.uniform(0, 1, ($N, 3)), fmt="%.17g")
PY
SNAP=tests/golden/cli/multi-simple-o0.snapshot
s=$(date +%s%N); oracle/_ref/interactive_emulator_ref interactive_mode $SNAP --quiet < /tmp/stream_pts.txt > /tmp/out_ref.txt; e=$(date +%s%N)
echo "reference CLI (CPU): $N points in $(( (e - s) / 1000000 )) ms"
s=$(date +%s%N); madaiemulator_b200/host/emub_interactive_emulator interactive_mode $SNAP --quiet < /tmp/stream_pts.txt > /tmp/out_gpu.txt; e=$(date +%s%N)
echo "streaming tool (GPU): $N points in $(( (e - s) / 1000000 )) ms"
if [ -x oracle/_ref/interactive_emulator_dropin_multi ]; then
  # the reference's own per-point loop (fscanf / emulate_point_multi / fprintf + fflush per point), every point on the GPU
  s=$(date +%s%N); oracle/_ref/interactive_emulator_dropin_multi interactive_mode $SNAP --quiet < /tmp/stream_pts.txt > /tmp/out_dropin.txt; e=$(date +%s%N)
  echo "reference CLI on the engine, point by point (few-points path): $N points in $(( (e - s) / 1000000 )) ms"
fi
python - <<PY
import numpy as np
a = np.loadtxt("/tmp/out_ref.txt"); b = np.loadtxt("/tmp/out_gpu.txt")
print("lines", a.size, b.size, "max abs diff", np.max(np.abs(a - b)))
PY
