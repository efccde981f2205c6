EMU=1 ROWS=65536 WHAT=field python scratch/cap_test2.py 2>&1 | grep -v Warning | tail -3
EMU=1 ROWS=65536 WHAT=ops python scratch/cap_test2.py 2>&1 | grep -v Warning | tail -8
