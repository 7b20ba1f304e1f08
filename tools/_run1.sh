P=$PWD/microstructure_fingerprinting_b200
for snr in 30 10 5; do
for lib in libmfb200_prev.so libmfb200.so; do
  MFB_LIB=$P/$lib timeout 300 python tools/quick_fit_bench.py 32768 0.3 0 $snr 2>&1 | tail -1
done; done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
