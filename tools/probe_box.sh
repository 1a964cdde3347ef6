#!/bin/bash
# Probe of the GPU box for what a stock PrMers (OpenCL + GMP) CPU baseline would need.
echo "== nproc: $(nproc)"; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" 
echo "== /etc/OpenCL/vendors:"; ls -la /etc/OpenCL/vendors 2>&1
echo "== clinfo:"; which clinfo 2>&1; clinfo -l 2>&1 | head -5
echo "== pocl:"; ls /usr/lib/x86_64-linux-gnu | grep -i -E "pocl|OpenCL" 2>&1; find / -xdev \( -name "libpocl*" -o -name "pocl.icd" \) 2>/dev/null | head
echo "== CL headers:"; find / -xdev -name "cl.h" -path "*CL*" 2>/dev/null | head
echo "== gmp:"; find / -xdev \( -name "gmp.h" -o -name "gmpxx.h" -o -name "libgmp*" \) 2>/dev/null | head; pkg-config --modversion gmp 2>&1
echo "== toolchains:"; for t in go javac node rustc clang cmake; do printf "%s: " $t; which $t 2>&1 || echo absent; done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
