"""Static SASS instruction mix of selected kernels in libh2o_b200.so."""
import re, collections, subprocess, sys
so = 'silver2_isaacsim_b200/lib/libh2o_b200.so'
pats = sys.argv[1:] or ['tile_kernelIfLi0ELi1ELb0ELb0', 'tile_kernelIdLi0ELi1ELb0ELb0', 'tile_kernelIfLi0ELi0ELb1ELb0']
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
for p in re.split(r'\n\s+Function : ', txt)[1:]:
    name = p.split('\n', 1)[0]
    if any(x in name for x in pats):
        ins = re.findall(r'^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P[0-9T]+ )?([A-Z0-9_]+)', p, re.M)
        c = collections.Counter(ins)
        print(name, 'total', len(ins))
        print('  ', ', '.join(f'{k}:{v}' for k, v in c.most_common(45)))
