"""Which Blackwell-native instructions each hot kernel of libp2vit_b200.so contains (cuobjdump -sass):
UTCIMMA = tcgen05.mma kind::i8, LDTM / STTM = tcgen05.ld / .st, UTMALDG / UTMASTG = TMA loads / stores, UTCBAR =
tcgen05.commit, IMMA = the legacy mma.sync path, IDP.4A = dp4a (the Swin window attention kernel's products).

    python tools/sass_evidence.py > profiles/r2_sass_evidence.txt
"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'diff_vit_b200', 'libp2vit_b200.so')
PAT = re.compile(r'\b(UTCIMMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|IMMA|IDP|UBLKCP|SYNCS)[A-Z0-9_.x]*')
out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
kernels = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name).replace('void p2v::', '')
        kernels[name] = [0, collections.Counter()]
        continue
    if name and re.search(r'/\*[0-9a-f]{4,}\*/', line):
        kernels[name][0] += 1
        m = PAT.search(line)
        if m:
            kernels[name][1][m.group(0)] += 1
print('kernel: SASS instructions | Blackwell-native mnemonics (count)')
for k, (n, c) in kernels.items():
    if any(s in k for s in ('attention', 'gemm_i8_tc', 'gemm_i8_bs', 'layernorm_int_pot')):
        tags = ', '.join('%s x%d' % kv for kv in sorted(c.items()) if not kv[0].startswith('SYNCS')) or '(none)'
        print('%-44s %6d | %s' % (k, n, tags))
