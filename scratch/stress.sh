for i in 1 2 3 4; do python scratch/dbg_nn2.py 2>&1 | grep -E "^[0-9]" | python -c "
import sys,re
for l in sys.stdin:
    v=[float(x) for x in re.findall(r\"'([0-9.e+-]+)'\", l)]
    print(l[:14], 'bad', sum(1 for x in v if x>1e-4), 'of', len(v))
"; done
