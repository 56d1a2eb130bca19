for s in 1 2 4; do SDPB_Q2_SHARE=1 SDPB_Q2_SPLIT=$s python tests/q2m_worker.py 2>&1 | tail -3; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
