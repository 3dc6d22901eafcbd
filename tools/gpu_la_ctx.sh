# la_ctx2 check with tight timeouts: tests, then stand-alone timing with the v2 and v1 context pass
mkdir -p gpurun_out; P=gpurun_out/${1:-lac}
timeout 150 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "linattn" > ${P}_pytest.log 2>&1; echo "pytest exit=$?"; tail -4 ${P}_pytest.log
timeout 60 python tools/run_linattn.py 2>&1 | tee ${P}_plain.log
IDIFF_LA_CTX_V2=1 timeout 60 python tools/run_linattn.py 2>&1 | tee ${P}_plain_ctxv1.log
