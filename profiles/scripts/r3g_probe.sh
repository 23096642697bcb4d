set -x
for lib in libgymwipe_b200.so variants/lib_macronj.so; do
echo LIB $lib; GYMWIPE_B200_LIB=gymwipe_b200/lib/$lib timeout 300 python profiles/scripts/cfg4_profile.py 64 2>&1 | tail -1
done
GYMWIPE_B200_LIB=gymwipe_b200/lib/variants/lib_macronj.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
