"""Mnemonic histogram of the SASS of every kernel in the built library (cuobjdump -sass):
   python scripts/sass_histogram.py [lib] > profiles/r2_sass_histogram.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'nxblink_b200', 'libnxblink_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        kern = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
        hist[kern] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)', line)
    if m and kern:
        hist[kern][m.group(1).split('.')[0]] += 1
KEYS = ('DMMA', 'HMMA', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'LDGSTS', 'LDS', 'STS', 'LDG', 'STG', 'ATOMS', 'ATOMG', 'RED', 'CCTL',
        'MUFU', 'DFMA', 'DMUL', 'DADD', 'FFMA', 'FMUL', 'FADD', 'IMAD', 'LOP3', 'SHFL', 'BAR', 'VOTE', 'BRA', 'BSSY', 'BSYNC', 'CALL')
print('SASS mnemonic counts per kernel of', os.path.basename(lib), '(sm_100a)')
print('(bulk-async mnemonics: UBLKCP = cp.async.bulk, the shared -> global copies of the new blink lists in blink_kernel; UTMALDG / SYNCS = tensor-map loads / mbarrier, not used: see DESIGN.md section 4)')
for k, h in hist.items():
    tot = sum(h.values())
    if tot < 200:
        continue
    print('\n%s\n  total %d' % (k[:150], tot))
    print('  ' + '  '.join('%s %d' % (m, h.get(m, 0)) for m in KEYS if h.get(m, 0) or m in ('DMMA', 'UBLKCP', 'UTMALDG', 'SYNCS')))
