"""Dev tool: per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (profiles/rNN_sass_summary.txt).
usage: cuobjdump -sass pcgcv1_b200/libpcgc_b200.so | python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, re, subprocess, sys
cur = None
counts = collections.OrderedDict()
keys = ['UTCHMMA', 'UTMALDG', 'UBLKCP', 'LDTM', 'UTCBAR', 'UTMASTG', 'HMMA', 'CREDUX', 'LDGSTS', 'ELECT']
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    if cur:
        for k in keys:
            if re.search(r'\b' + k + r'\b', line):
                counts[cur][k] += 1
names = subprocess.run(['c++filt'] + list(counts.keys()), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic counts per kernel of pcgcv1_b200/libpcgc_b200.so (cuobjdump -sass, sm_100a).")
print("# UTCHMMA = tcgen05.mma kind::f16, UTMALDG = cp.async.bulk.tensor (TMA load), UBLKCP = cp.async.bulk, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,")
print("# CREDUX = redux.sync (GPU range decoder), LDGSTS = cp.async, ELECT = elect.sync.  HMMA (legacy mma.sync) must not appear.")
tot = collections.Counter()
for (k, c), n in zip(counts.items(), names):
    if sum(c.values()) == 0:
        continue
    short = re.sub(r'pcgc::\(anonymous namespace\)::', '', n)
    short = re.sub(r'\(.*', '', short).replace('void ', '')
    print("%-64s %s" % (short[:64], " ".join("%s=%d" % (a, b) for a, b in sorted(c.items()))))
    tot.update(c)
print("TOTAL", " ".join("%s=%d" % (a, b) for a, b in sorted(tot.items())))
