#!/bin/sh
# build variants of the library with different -D knobs and run tools/bench_v2.py on each: tools/v2_ab.sh "<bench args>" name:-Dflag ...
args=$1; shift
mkdir -p gpurun_out/variants
for v in "$@"; do
  name=${v%%:*}; flags=$(echo "${v#*:}" | tr ',' ' ')
  python - "$name" $flags <<'PY'
import subprocess, sys
sys.path.insert(0, '.')
from wab_gym_b200 import build as b
name, flags = sys.argv[1], [f for f in sys.argv[2:] if f != '-']
cmd = [b.find_nvcc()] + b.NVCC_FLAGS + flags + ['-o', 'gpurun_out/variants/libwab_%s.so' % name] + b.SOURCES
r = subprocess.run(cmd, capture_output=True, text=True)
if r.returncode: print(name, 'build failed', r.stderr[-800:])
PY
  WAB_LIB=gpurun_out/variants/libwab_$name.so python tools/bench_v2.py $args | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', '%.4g turns/s  %.3f ms/turn  frac %.3f' % (d['world_turns_per_s'], d['ms_per_turn'], d['frac']))"
done
