ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:oda_pair -c 3 \
    -o gpurun_out/oda_pair -f python tools/ncu_step.py --model ODA > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
