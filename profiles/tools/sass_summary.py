"""Counts of the instructions that prove the hardware paths (TMA: UTMALDG / UTMASTG / UTMAPF, FP64 tensor pipe: DMMA, mbarrier:
SYNCS, cp.async: LDGSTS, REDUX) per kernel, from cuobjdump -sass of the built objects.  python profiles/tools/sass_summary.py"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, 'qio_b200', 'csrc')
KEY = re.compile(r'^(UTMALDG|UTMASTG|UTMAPF|UTMACMDFLUSH|DMMA|SYNCS|LDGSTS|REDUX)')
print("# SASS evidence: cuobjdump -sass of the objects built by `make -C qio_b200/csrc` (nvcc 12.9, sm_100a)")
for obj in ('rotate_tma.o', 'prep.o', 'sweep3.o', 'rotate.o', 'pairs.o', 'sweep2.o'):
    out = subprocess.run(['cuobjdump', '-sass', os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    fn = None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r'\(.*', '', fn).replace('void ', '')
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)', line)
        if m and KEY.match(m.group(1)):
            counts.setdefault(fn, collections.Counter())[m.group(1)] += 1
    print(f'\n== {obj}')
    for fn, c in counts.items():
        print(f'  {fn}: ' + ', '.join(f'{k} x{v}' for k, v in sorted(c.items())))
