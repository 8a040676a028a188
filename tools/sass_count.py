"""Static SASS instruction counts per kernel (cuobjdump -sass on the built objects): total, IMAD*, IADD3/LOP3/SHF/SEL,
LDG/STG.  Usage: python tools/sass_count.py <regex on the demangled name> [obj ...]"""
import glob, re, subprocess, sys
pat = sys.argv[1] if len(sys.argv) > 1 else ''
objs = sys.argv[2:] or sorted(glob.glob('mpyc_b200/csrc/_obj/*.o'))
for obj in objs:
    out = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
    for blk in out.split('Function : ')[1:]:
        name = blk.split('\n', 1)[0].strip()
        dem = subprocess.run(['cu++filt', name], capture_output=True, text=True).stdout.strip().replace('void ', '')
        dem = dem[:dem.index('>(') + 1] if '>(' in dem else dem.split('(')[0]
        dem = dem.replace('(int)', '').replace('(bool)', '')
        if not re.search(pat, dem):
            continue
        ins = re.findall(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', blk)
        def cnt(rx):
            return sum(1 for i in ins if re.match(rx, i))
        print(f"{dem:48s} total={len(ins):5d} IMAD.WIDE={cnt(r'IMAD\.WIDE'):4d} IMAD*={cnt(r'IMAD'):4d} "
              f"IADD3={cnt(r'IADD3|IADD'):4d} LOP/SHF/SEL={cnt(r'LOP3|SHF|SEL'):4d} LDG={cnt(r'LDG'):3d} STG={cnt(r'STG'):3d}")
