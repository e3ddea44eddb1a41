ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02.csv python tools/profile_step.py > gpurun_out/p1.log 2>&1; tail -1 gpurun_out/p1.log
ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_step_r02 -f python tools/profile_step.py > gpurun_out/p2.log 2>&1; tail -1 gpurun_out/p2.log
ls -la gpurun_out/prof_step_r02.ncu-rep gpurun_out/launches_r02.csv
