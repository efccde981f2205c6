EMU=1 ROWS=65536 WHAT=ops python scratch/cap_test2.py 2>&1 | grep -v Warning | tail -5
EMU=0 ROWS=65536 WHAT=field python scratch/cap_test2.py 2>&1 | grep -v Warning | tail -3
EMU=1 ROWS=4096 WHAT=field python scratch/cap_test2.py 2>&1 | grep -v Warning | tail -3
