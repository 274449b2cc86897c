python tools/time_attn.py > gpurun_out/attn_fast.txt 2>&1; cat gpurun_out/attn_fast.txt
cp vit_exp_b200/libctk.so /tmp/libctk_fast.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -DCTK_NO_FAST_EXP2 -c vit_exp_b200/csrc/attention.cu -o /tmp/attention_slow.o && nvcc -shared -o vit_exp_b200/libctk.so vit_exp_b200/csrc/obj/core.o vit_exp_b200/csrc/obj/cpb.o vit_exp_b200/csrc/obj/elementwise.o vit_exp_b200/csrc/obj/gemm_tcgen05.o vit_exp_b200/csrc/obj/head.o /tmp/attention_slow.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo "--- exp2f build"; python tools/time_attn.py
cp /tmp/libctk_fast.so vit_exp_b200/libctk.so
