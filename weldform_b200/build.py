"""Build libweldform_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc. Cross-compiles without a GPU."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libweldform_b200.so")
SOURCES = ["wf_ctx.cu", "wf_neighbour.cu", "wf_pair.cu", "wf_pair_fm.cu", "wf_step.cu", "wf_comm.cu"]
# wf_pair_fm.cu holds the contracted-arithmetic variants of the pair kernels: the one translation unit compiled with FMA contraction
FMAD_SOURCES = {"wf_pair_fm.cu"}
# -fmad=false: the pair terms must round exactly like the reference's x86-64 build (no FMA contraction); see DESIGN.md
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Wno-deprecated-declarations", "-diag-suppress", "177,550"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "weldform_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=()):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        flags = [("-fmad=true" if (f == "-fmad=false" and src in FMAD_SOURCES) else f) for f in NVCC_FLAGS]
        cmd = [_nvcc()] + flags + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [_nvcc(), "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    return LIB


HOST_DIR = os.path.join(HERE, "host")
HOST_BIN = os.path.join(HOST_DIR, "bin")
EXAMPLES = os.path.join(os.path.dirname(HERE), "examples")


def build_host_examples(force=False):
    """Compile examples/*.cpp (drivers written against the reference-style Domain API in weldform_b200/host/) and link them
    with libweldform_b200.so. Host C++ only -- g++, no nvcc."""
    build()
    os.makedirs(HOST_BIN, exist_ok=True)
    out = []
    for src in sorted(os.listdir(EXAMPLES)):
        if not src.endswith(".cpp"):
            continue
        exe = os.path.join(HOST_BIN, src[:-4])
        deps = [os.path.join(EXAMPLES, src), LIB] + [os.path.join(HOST_DIR, f) for f in os.listdir(HOST_DIR) if f.endswith((".h", ".cpp"))]
        if force or not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
            cmd = ["g++", "-std=c++14", "-O2", "-fpermissive", "-w", "-fopenmp", "-I" + HOST_DIR, "-I" + os.path.join(os.path.dirname(HERE), "include"),
                   os.path.join(EXAMPLES, src), "-o", exe, "-L" + HERE, "-lweldform_b200", "-Wl,-rpath,$ORIGIN/../.."]
            subprocess.check_call(cmd)
        out.append(exe)
    return out


REF_EXAMPLES = "/root/reference/examples"
REF_DRIVERS = ["Compression", "Compression_thermal", "Contact_Compression_1010", "FSW",
               "Compression_1010", "Compression_1010_half_ghost", "CompressionVel", "Compression_TRIP_vel", "Contact_Compression_half_ghost",
               "Contact_Compression_TRIP_2surf", "Contact_Compression_1010_1side", "Contact_Compression_2surf_short", "Contact_Compression_vel",
               "Contact_Compression_1010_unif", "Contact_Compression_thermal", "Contact_Compression_Rot", "FSW_thin",
               "slice_test", "Traction", "Traction_JC", "Traction_Hollomon", "Twisting"]


def build_reference_drivers(force=False):
    """The reference's OWN example drivers (BASELINE.json's configurations), compiled UNCHANGED from /root/reference/examples against the host
    mirror (weldform_b200/host/) and linked with libweldform_b200.so -> host/bin/ref_<name>. Only where /root/reference exists (the build
    container); the binaries are git-ignored and travel to the GPU box like the library. No reference source is copied."""
    if not os.path.isdir(REF_EXAMPLES):
        return [os.path.join(HOST_BIN, "ref_" + d) for d in REF_DRIVERS if os.path.exists(os.path.join(HOST_BIN, "ref_" + d))]
    build()
    os.makedirs(HOST_BIN, exist_ok=True)
    out, procs = [], []
    for d in REF_DRIVERS:
        src = os.path.join(REF_EXAMPLES, d + ".cpp")
        exe = os.path.join(HOST_BIN, "ref_" + d)
        deps = [src, LIB] + [os.path.join(HOST_DIR, f) for f in os.listdir(HOST_DIR) if f.endswith((".h", ".cpp"))]
        if force or not os.path.exists(exe) or any(os.path.getmtime(x) > os.path.getmtime(exe) for x in deps):
            cmd = ["g++", "-std=c++14", "-O2", "-fpermissive", "-w", "-fopenmp", "-D_USE_MATH_DEFINES", "-I" + HOST_DIR, "-I" + os.path.join(os.path.dirname(HERE), "include"),
                   src, "-o", exe, "-L" + HERE, "-lweldform_b200", "-Wl,-rpath,$ORIGIN/../.."]
            procs.append((d, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        out.append(exe)
    for d, p in procs:
        o, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(o)
            raise RuntimeError("reference driver %s does not compile against the host mirror" % d)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
    for exe in build_host_examples() + build_reference_drivers():
        print(exe)
