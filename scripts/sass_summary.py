"""Per-kernel SASS mnemonic counts of libgwn.so (cuobjdump -sass): evidence that the kernels are Blackwell-native.
usage: python scripts/sass_summary.py [lib] > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), '..', 'multimodal_outage_b200', 'libgwn.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
keys = ['UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'UBLKCP', 'UBLKPF', 'SYNCS', 'REDG', 'MUFU.TANH']
print('Per-kernel SASS mnemonic counts of the final round-2 libgwn.so (cuobjdump -sass, sm_100a): UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor\n'
      'load / store, LDTM / STTM = tcgen05.ld / st (TMEM), UBLKCP / UBLKPF = cp.async.bulk (copy / L2 prefetch), SYNCS = mbarrier ops, REDG = red.global.\n')
tot = collections.Counter()
for name, body in zip(names, re.split(r'Function : \S+', sass)[1:]):
    ins = re.findall(r'^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', body, flags=re.M)
    c = collections.Counter()
    for i in ins:
        for k in keys:
            if i.startswith(k):
                c[k] += 1
    tot.update(c)
    short = re.sub(r'\(.*', '', name)
    print(short)
    print(f'    instructions {len(ins)}  ' + '  '.join(f'{k} {c[k]}' for k in keys if c[k]))
print('\nlibrary totals: ' + '  '.join(f'{k} {tot[k]}' for k in keys))
