PRELOAD=0 python scratch/emu_test.py 2>&1 | tail -5
PRELOAD=1 python scratch/emu_test.py 2>&1 | tail -5
PRELOAD=1 CUBLAS_EMULATE_SINGLE_PRECISION=1 python scratch/emu_test.py 2>&1 | tail -5
PRELOAD=1 CUBLAS_EMULATE_SINGLE_PRECISION=1 CUBLAS_EMULATION_STRATEGY=eager python scratch/emu_test.py 2>&1 | tail -5
