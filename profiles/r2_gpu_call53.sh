set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cnn.py "tests/test_gpu_model.py::test_constructor_surface_and_variable_scopes" -q -x > gpurun_out/s53_pytest_cnn.txt 2>&1
timeout 300 python bench.py --cnn --no-convergence > gpurun_out/s53_bench_cnn.json 2> gpurun_out/s53_bench_cnn.err
