# A/B two builds of the library in the same gpurun call: tools/ab.sh <rounds>
for i in $(seq 1 ${1:-2}); do
for lib in gpt_b200/lib/prev.so gpt_b200/lib/libcgpt_b200.so; do
GPT_B200_LIBRARY=$PWD/$lib timeout 300 python bench.py --no-e2e --no-cpu --no-cg --steps 300 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('AB', '$lib', round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
GPT_B200_LIBRARY=$PWD/$lib python tools/cg_bench.py | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ABCG', '$lib', round(d['ms_per_iteration'],3))"
done; done
