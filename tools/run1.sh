N="ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
timeout 600 $N -k regex:'k_conv_umma<.int.4, .int.32, .int.3>' -s 1 -c 1 -o gpurun_out/prof_conv4323 -f python tests/gpu_ncu_step.py 3840 learner > gpurun_out/ncu1.log 2>&1
timeout 600 $N -k regex:'k_wgrad_umma<.int.4, .int.32>' -s 1 -c 1 -o gpurun_out/prof_wgrad432 -f python tests/gpu_ncu_step.py 3840 learner > gpurun_out/ncu2.log 2>&1
timeout 600 $N -k regex:'k_conv_umma<.int.2, .int.16, .int.3>' -s 1 -c 1 -o gpurun_out/prof_conv2163 -f python tests/gpu_ncu_step.py 3840 learner > gpurun_out/ncu3.log 2>&1
tail -n 2 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log; ls -la gpurun_out/*.ncu-rep
