for cfg in "DDPM_SPLITK=1 DDPM_GN_SOLO_ELEMS=65536" "DDPM_SPLITK=0 DDPM_GN_SOLO_ELEMS=65536" "DDPM_SPLITK=1 DDPM_GN_SOLO_ELEMS=0" "DDPM_SPLITK=0 DDPM_GN_SOLO_ELEMS=0"; do
  env $cfg python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab.log 2> gpurun_out/ab.err
  python -c "
import json
l=json.loads(open('gpurun_out/ab.log').read().strip().splitlines()[-1])
print('$cfg', 'train ms', l['ms_per_step'], 'sampling ms', l['sampling']['ms_per_step_device_noise'])"
done
