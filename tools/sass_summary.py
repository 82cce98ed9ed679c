#!/usr/bin/env python
"""Opcode histogram of the hot kernels from the built objects (cuobjdump -sass), the evidence for DMMA / bulk-TMA / cp.async use.
   python tools/sass_summary.py > profiles/RNN_sass.txt      (run after build.sh; no GPU needed)"""
import collections, re, subprocess, sys
KERNELS = [("build/k1_expand.o", "k1_expand_pipe_kernel"), ("build/k1_expand.o", "k1_expand_kernel"), ("build/k1_expand.o", "k1_expand_slab_kernel"),
           ("build/k2_shell.o", "k2_quad_planar_vm_kernel"), ("build/k2_shell.o", "k2_quad_flat_vm_kernel"),
           ("build/k2_shell.o", "k2_shell_vm_kernel"), ("build/k2_solid.o", "k2_tet10_steplane_vm_kernel"),
           ("build/k2_solid.o", "k2_tet10_affine_vm_kernel"), ("build/k2_hex20.o", "k2_hex20_steplane_vm_kernel"),
           ("build/k2_hex20.o", "k2_bigsolid_grad_vm_kernel"), ("build/io_rdb.o", "record_points_dmma_kernel"),
           ("build/k3_fatigue.o", "k3_stream_kernel"), ("build/k3_gage.o", "gage_post_kernel")]
WATCH = ("DMMA", "UBLKCP", "LDGSTS", "SYNCS", "DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "BAR", "LDL", "STL")
cache = {}
for obj, name in KERNELS:
    if obj not in cache:
        cache[obj] = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", cache[obj])[1:]
    for f in funcs:
        mangled = f.split("\n", 1)[0].strip()
        if name not in mangled:
            continue
        ops = collections.Counter()
        for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*)(\.[A-Z0-9_.]+)?", f):
            ops[m.group(1)] += 1
            if m.group(1) == "DMMA" and m.group(2):
                ops["DMMA" + m.group(2)] += 1
        total = sum(v for k, v in ops.items() if "." not in k)
        demangled = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip().split("(")[0]
        print(f"{demangled}\n   instructions {total}: " + ", ".join(f"{k} {ops[k]}" for k in list(WATCH) + [k for k in ops if k.startswith('DMMA.')] if ops.get(k)))
