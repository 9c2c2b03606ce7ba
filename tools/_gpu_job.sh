python tools/lpt_probe.py 8 2>&1 | tail -19
python tools/lpt_probe.py 4 2>&1 | grep slowest
python tools/lpt_probe.py 2 2>&1 | grep slowest
