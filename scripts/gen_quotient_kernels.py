#!/usr/bin/env python
"""Emit battlezips-halo2_b200/csrc/gen_quotient.cu: h(X) of the Shot and Board circuits as straight-line sm_100a code.

For every known circuit and evaluation tier the library's own compiler (csrc/evalprog.h through bz_quotient_program -- host
only, no GPU) produces the stack program the interpreter `eval_program_kernel` would run; this script simulates its stack
statically and prints one kernel per program in which every stack slot, temporary and the accumulator is a named local
variable: registers instead of the interpreter's local-memory stack, no instruction fetch / decode / dispatch, rotations and
column offsets as literals.  A proving key finds its kernel by the FNV-1a hash of the program; any other circuit runs the
interpreter.  Same field operations in the same order, so the proof bytes are identical
(tests/test_gpu_prover.py::test_generated_quotient_matches_interpreter).

Run after changing the circuits or the compiler:   python scripts/gen_quotient_kernels.py
tests/test_evalprog_host.py::test_generated_quotient_kernels_are_current fails when the committed file is stale."""
import ctypes, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "battlezips-halo2_b200", "csrc", "gen_quotient.cu")
OPS = ["PUSH_P", "PUSH_S", "PUSH_C", "ADD", "SUB", "MUL", "NEG", "MULC", "ADDC", "FOLD", "STORE", "END", "MUL_T_STORE", "ACC_MULC", "TEE", "PUSH_T"]
VK_REPR = 0x1234567890ABCDEF1234567890ABCDEF


def programs(cs, k):
    """[(tier, code words, rotation table, hash)] of a constraint system (the programs do not depend on k beyond the rotation scale)."""
    import battlezips_halo2_b200 as bz
    from battlezips_halo2_b200.plonk import prover as PR
    lib = bz.load_library()
    circ, keep = PR.flatten_circuit(cs.to_ir(), k, VK_REPR)
    out = []
    for tier in range(3):
        n_code, n_rot, h = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint64()
        rc = lib.bz_quotient_program(ctypes.byref(circ), tier, None, 0, ctypes.byref(n_code), None, 0, ctypes.byref(n_rot), ctypes.byref(h))
        assert rc == 0, rc
        if not n_code.value:
            continue
        code = (ctypes.c_uint32 * n_code.value)()
        rot = (ctypes.c_int32 * max(1, n_rot.value))()
        rc = lib.bz_quotient_program(ctypes.byref(circ), tier, code, n_code.value, ctypes.byref(n_code), rot, n_rot.value, ctypes.byref(n_rot), ctypes.byref(h))
        assert rc == 0, rc
        out.append((tier, list(code), list(rot)[: n_rot.value], h.value))
    return out


def emit_kernel(name, code, rot):
    """Straight-line body for one program.  Stack slot i -> variable s<i>; temporaries t<j>; accumulator acc."""
    L = []
    sp, max_sp, tmps = 0, 0, set()
    for ins in code:
        op, x, y, c = OPS[ins & 15], (ins >> 4) & 0xfff, ins >> 16, ins >> 4
        if op in ("PUSH_P", "PUSH_S"):
            r = rot[y]
            idx = "jp" if r == 0 else f"((jp + {r & 0xffffffff}u) & mask)"
            base = "pb" if op == "PUSH_P" else "a.sbase"
            L.append(f"  s{sp} = fe_load({base} + ((uint64_t){x}u << a.logN) + {idx});")
            sp += 1
        elif op == "PUSH_C":
            L.append(f"  s{sp} = fe_load(cb + {c});"); sp += 1
        elif op in ("ADD", "SUB", "MUL"):
            sp -= 1
            L.append(f"  s{sp - 1} = fe_{op.lower()}(s{sp - 1}, s{sp});")
        elif op == "NEG":
            L.append(f"  s{sp - 1} = fe_neg(s{sp - 1});")
        elif op == "MULC":
            L.append(f"  s{sp - 1} = fe_mul(s{sp - 1}, fe_load(cb + {c}));")
        elif op == "ADDC":
            L.append(f"  s{sp - 1} = fe_add(s{sp - 1}, fe_load(cb + {c}));")
        elif op == "FOLD":
            sp -= 1
            L.append(f"  acc = fe_add(fe_mul(acc, fe_load(cb + {c})), s{sp});")
        elif op == "ACC_MULC":
            L.append(f"  acc = fe_mul(acc, fe_load(cb + {c}));")
        elif op == "MUL_T_STORE":
            L.append("  fe_store(a.out + (uint64_t)b * a.ostride + i, fe_mul(acc, fe_load(a.tev + (jp & (a.tn - 1)))));")
        elif op == "STORE":
            sp -= 1
            L.append(f"  fe_store(a.out + (uint64_t)b * a.ostride + ((uint64_t){c}u << a.logN) + i, s{sp});")
        elif op == "TEE":
            tmps.add(c); L.append(f"  t{c} = s{sp - 1};")
        elif op == "PUSH_T":
            L.append(f"  s{sp} = t{c};"); sp += 1
        else:
            raise ValueError(op)
        max_sp = max(max_sp, sp)
        assert sp >= 0
    assert sp == 0, "program leaves values on the stack"
    decl = ", ".join([f"s{i}" for i in range(max_sp)] + [f"t{j}" for j in sorted(tmps)])
    head = [f"__global__ void __launch_bounds__(128) {name}(const __grid_constant__ EvalArgs<FpP> a) {{",
            "  const uint32_t N = 1u << a.logN, mask = N - 1;",
            "  const uint32_t i = a.first + blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;",
            "  if (i >= (N >> a.stride_log)) return;",
            "  const uint32_t jp = i << a.stride_log;",
            "  const Fe<FpP>* pb = a.pbase + (uint64_t)b * a.pstride;",
            "  const Fe<FpP>* cb = a.consts + (uint64_t)b * a.cstride;",
            f"  Fe<FpP> {decl};" if decl else "",
            "  Fe<FpP> acc = fe_zero<FpP>();"]
    return "\n".join(head + L + ["}"])


def render():
    from battlezips_halo2_b200.circuits import shot_circuit, board_circuit
    entries, kernels, seen = [], [], set()
    for cname, make, k in (("shot", shot_circuit, 11), ("board", board_circuit, 12)):
        cs = make(0)[0]
        for tier, code, rot, h in programs(cs, k):
            if h in seen:
                continue
            seen.add(h)
            name = f"q_{cname}_t{tier}"
            kernels.append(f"// {cname} circuit, tier {tier}: {len(code)} program words, {sum(1 for w in code if OPS[w & 15] in ('MUL', 'MULC', 'FOLD', 'ACC_MULC', 'MUL_T_STORE'))} multiplications per point\n" + emit_kernel(name, code, rot))
            entries.append((h, name))
    out = ["// GENERATED by scripts/gen_quotient_kernels.py -- do not edit.  h(X) of the known circuits as straight-line code: one kernel per",
           "// (circuit, evaluation tier), selected by the hash of the program the library compiles for a proving key (prover.cu",
           "// `upload_programs`); every other circuit runs the interpreter (poly.cuh `eval_program_kernel`).",
           '#include "prover_impl.h"', "namespace bz {", ""]
    out += [k + "\n" for k in kernels]
    for h, name in entries:
        out.append(f"static void launch_{name}(const EvalArgs<FpP>& a, dim3 grid, cudaStream_t st) {{ {name}<<<grid, 128, 0, st>>>(a); }}")
    out.append("")
    out.append("QuotientLaunchFn find_generated_quotient(uint64_t hash) {")
    out.append("  switch (hash) {")
    for h, name in entries:
        out.append(f"    case 0x{h:016x}ull: return launch_{name};")
    out.append("    default: return nullptr;")
    out.append("  }")
    out.append("}")
    out.append("")
    out.append("}  // namespace bz")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    text = render()
    if "--check" in sys.argv:
        sys.exit(0 if os.path.exists(OUT) and open(OUT).read() == text else 1)
    open(OUT, "w").write(text)
    print("wrote", OUT, len(text), "bytes")
