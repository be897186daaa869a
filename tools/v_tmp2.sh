for wb in 544 768 1024 1536; do echo "WB $wb"; FDL_JPEG_WINDOW_BITS=$wb python tools/jpeg_bench.py 256 3 90 2>&1 | grep -E "phases|frames/s"; done
