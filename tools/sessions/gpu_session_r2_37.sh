python tools/grey_microbench.py 2>&1 | tail -1
for v in variants/*.so; do echo $v; HHSR_LIB=$PWD/$v python tools/grey_microbench.py 2>&1 | tail -1; done
