for i in 1 2; do
for lib in gym_art_b200/_exp/libquadsim_f4.so gym_art_b200/libquadsim.so; do
  QUADSIM_LIB=$PWD/$lib python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', 'us/step %.3f' % (d['ms_per_step']*1e3), 'frac %.3f' % d['roofline']['frac'], 'idle %.3f' % (d['ms_per_step_from_idle']*1e3), 'l2res %.3f' % (d['l2_resident_same_batch']['ms_per_step']*1e3), d['clocks'])"
done; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu --format=csv
