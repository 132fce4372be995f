bash tools/profile_round.sh r2
( echo "# compute-sanitizer on the torch-free ABI driver (tools/abi_driver.cu), r2 final code"; echo; echo "## memcheck"; timeout 900 compute-sanitizer --tool memcheck tools/abi_driver 2>&1 | tail -16; echo; echo "## synccheck"; timeout 600 compute-sanitizer --tool synccheck tools/abi_driver 2>&1 | tail -4 ) > gpurun_out/sanitizer_r2.txt
tail -30 gpurun_out/sanitizer_r2.txt
