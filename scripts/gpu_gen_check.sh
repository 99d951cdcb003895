set -x
python -m pytest tests/test_gpu_units.py -x -q -m gpu -k "generat" 2>&1 | tail -15
