#!/bin/sh
# Commit-able SASS listing of the hot kernels (profiles/<tag>_sass_*.txt) from the in-tree library.
set -e
cd "$(dirname "$0")/.."
tag=${1:-r1}
cuobjdump -sass wab_gym_b200/libwab_b200.so > /tmp/wab_all.sass
python - "$tag" <<'PY'
import re, sys
tag = sys.argv[1]
txt = open('/tmp/wab_all.sass').read()
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n', 1)[0]
    for key, out in (('wab_step_kernelILb0ELi1ELi24E', 'step_lpe1'), ('wab_step_kernelILb0ELi16E', 'step_lpe16'), ('wab2_turn_kernel', 'v2_turn')):
        if key in name:
            body = 'Function : ' + f
            n = len(re.findall(r'^\s+/\*[0-9a-f]{4,5}\*/', body, flags=re.M))
            open('profiles/%s_sass_%s.txt' % (tag, out), 'w').write('# %d SASS instructions (%.1f KB), cuobjdump -sass wab_gym_b200/libwab_b200.so\n' % (n, n * 16 / 1024) + body)
            print(out, n)
PY
