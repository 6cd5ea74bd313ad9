python -m pytest tests/test_gpu_quantize_tc.py -x -q 2>&1 | tail -3
python profiles/exp_qtc.py
python profiles/quantizer_phases.py
