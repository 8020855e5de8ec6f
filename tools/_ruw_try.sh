set -u
for i in 1 2; do
timeout 600 python tools/phase_times.py 2>&1 | grep ae_decode
ECHO_CONV_EPILOGUE=0 timeout 600 python tools/phase_times.py 2>&1 | grep ae_decode
done
