CUBLAS_EMULATE_SINGLE_PRECISION=1 python scratch/emu_test2.py 2>&1 | tail -22
