for g in 32 64 128; do
  PF_L2_FETCH_GRANULARITY=$g python bench.py --no-cpu-baseline --steps 3 --warmup 3 2> gpurun_out/gran_$g.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('gran',$g,'lookup ms',d['ms_lookup_kernel'],'gather GB/s',d['roofline_lookup']['random_sector_gather_gbs'],'align ms',d['ms_align_pipeline'])
"
done
