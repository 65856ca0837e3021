"""Per-subroutine SASS comparison of two builds of the solve kernel (no GPU needed): the noinline device functions keep
their own labels inside the kernel's section, so a change that should not touch a hot function can be checked by
comparing that function's instruction stream.

    cuobjdump -xelf all old.o; nvdisasm -c *.cubin > old.txt      (same for new.o -> new.txt)
    python scripts/sass_compare.py old.txt new.txt

Used in round 1 after an unused call site in qp_solve_gi had cost 5 %: ptxas re-rolls MOV / IMAD.MOV
encodings between builds, so equal length + same opcode classes is the practical criterion."""
import re, collections, sys
def parse(path):
    res = collections.OrderedDict(); cur=None
    for line in open(path):
        m = re.match(r'^(\$?\S+):\s*$', line)
        if m and ('dgsqp_solve_kernel' in m.group(1)) and not m.group(1).startswith('.L'):
            name = m.group(1)
            # kernel variant + internal function name
            var = 'T' if 'kernelILb1' in name.split('$')[1 if name.startswith('$') else 0] else 'F'
            mm = re.search(r'\$_ZN43_INTERNAL_\w+?_GLOBAL__N__\w+?_cu_[0-9a-f]{8}(\d+)(\w+)', name)
            fn = mm.group(2)[:int(mm.group(1))] if mm else 'ENTRY'
            tail = name[-40:]
            cur = (var, fn, tail if fn!='ENTRY' else ''); res[cur]=[]; continue
        m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(.*?);', line)
        if m and cur is not None:
            ins = re.sub(r'0x[0-9a-f]+','IMM', m.group(1)); ins = re.sub(r'`\([^)]*\)','LBL',ins)
            res[cur].append(ins)
    return res
A = parse(sys.argv[1]); B = parse(sys.argv[2])
print(len(A), len(B))
ka = {(v,f):k for k in A for v,f,_ in [k]}
tot=0
for k in B:
    v,f,_ = k
    if v!='T': continue
    a = A.get(ka.get((v,f)))
    if a is None: print('NEW', v, f, len(B[k])); continue
    if a != B[k]:
        print('DIFF', v, f, len(a), len(B[k])); tot+=1
    else:
        pass
print('differing SM=true functions:', tot, 'of', sum(1 for k in B if k[0]=='T'))
