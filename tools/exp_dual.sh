python tools/exp_dual.py 4,4
python tools/exp_dual.py 3,5
