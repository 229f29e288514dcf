for v in 1; do for tile in 16,2,1,1; do for z in 32 16; do for st in 0 1; do for a in 0 2; do
CGPTB_STCS=$st CGPTB_ZSLAB=$z CGPTB_TILE=$tile CGPTB_DHOP_VARIANT=$v CGPTB_ABLATE=$a timeout 300 python bench.py --no-e2e --no-cpu --steps 200 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('variant',$v,'tile','$tile','zslab',$z,'stcs',$st,'ablate',$a,round(d['ms_per_step'],4),round(d['roofline']['frac'],3),d['clocks']['sm_mhz'],d['clocks']['reasons'])"
done; done; done; done; done
