# build variants of the library with different launch bounds / pipeline depths and time them
SRC=lss2_multimodal_nu_b200/csrc
for v in "4 4 3" "5 3 3" "6 2 3" "5 4 3" "4 4 4" "5 3 4"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared -DLSS_FWD_MINB=$1 -DLSS_FWD_STAGES=$2 -DLSS_BWD_MINB=$3 \
    -I include -I $SRC -o lss2_multimodal_nu_b200/liblss_b200.so $SRC/lss_abi.cu 2>&1 | grep -i "error"
  echo -n "fwd minb=$1 stages=$2 bwd minb=$3: "
  timeout 100 python bench.py --no-cpu-baseline --steps 200 --e2e-steps 4 --in-flight 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step']*1e3,1), d['roofline']['kernels_us'])"
done
