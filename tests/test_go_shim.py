"""The Go shim is committed as source (go/pkg/core/hnsw/gpu_cuda.go) but cannot be compiled here (no Go
toolchain).  What CAN be checked mechanically: every C symbol, constant and struct the cgo code names is declared
in include/kektordb_gpu.h, each call passes as many arguments as the declaration takes, both build-tag variants
define the same Go API, and the file follows the build-tag pattern of the reference's own cgo file
(pkg/core/distance/distance_rust.go:1-17)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO = os.path.join(ROOT, "go", "pkg", "core", "hnsw")


def _header():
    text = open(os.path.join(ROOT, "include", "kektordb_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(kdbgpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        decls[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    consts = set(re.findall(r"#define\s+(KDBGPU_[A-Z0-9_]+)", text))
    types = set(re.findall(r"typedef struct (?:\{[^}]*\}|kdbgpu_[a-z_]+)\s*(kdbgpu_[a-z_]+)\s*;", text, flags=re.S))
    return decls, consts, types


def _calls(src):
    """(name, argument count) of every C.kdbgpu_* call, with nested parentheses handled."""
    out = []
    for m in re.finditer(r"C\.(kdbgpu_[a-z0-9_]+)\(", src):
        i, depth, n_args, seen = m.end(), 1, 0, False
        while depth:
            c = src[i]
            if c == "(":
                depth += 1
            elif c == ")":
                depth -= 1
            elif c == "," and depth == 1:
                n_args += 1
            elif not c.isspace():
                seen = True
            if c == "/" and src[i:i + 2] == "/*":
                i = src.index("*/", i) + 1
            i += 1
        out.append((m.group(1), n_args + 1 if seen else 0))
    return out


def test_cgo_file_uses_only_declared_symbols_with_the_declared_arity():
    decls, consts, types = _header()
    src = open(os.path.join(GO, "gpu_cuda.go")).read()
    assert src.startswith("//go:build cuda\n")
    assert '#include "kektordb_gpu.h"' in src and "#cgo LDFLAGS: -lkektordb_gpu" in src and 'import "C"' in src
    calls = _calls(src)
    assert len(calls) >= 25
    for name, n_args in calls:
        assert name in decls, f"{name} is not declared in include/kektordb_gpu.h"
        assert decls[name] == n_args, f"{name}: called with {n_args} arguments, declared with {decls[name]}"
    for c in set(re.findall(r"C\.(KDBGPU_[A-Z0-9_]+)", src)):
        assert c in consts, c
    for t in set(re.findall(r"\*C\.(kdbgpu_[a-z_]+)\b(?!\()", src)):
        assert t in types, t
    # the asynchronous shape: submit / poll / take, and no blocking one-query call anywhere
    used = {n for n, _ in calls}
    assert {"kdbgpu_batcher_submit", "kdbgpu_batcher_poll", "kdbgpu_batcher_take", "kdbgpu_refresher_set_row",
            "kdbgpu_arena_load_dir", "kdbgpu_set_graph_file"} <= used
    assert "kdbgpu_batcher_search" not in used


def test_both_build_tag_variants_define_the_same_go_api():
    cuda = open(os.path.join(GO, "gpu_cuda.go")).read()
    stub = open(os.path.join(GO, "gpu_nocuda.go")).read()
    assert stub.startswith("//go:build !cuda\n")
    meth = lambda s: set(re.findall(r"^func \(h \*Index\) (\w+)\(", s, flags=re.M))
    assert meth(stub) <= meth(cuda)
    for name in ("AttachGPU", "DetachGPU", "GPUFlush", "searchWithScoresGPU", "gpuNoteAdd", "gpuNoteRow", "gpuNoteDelete",
                 "gpuNoteRemove", "gpuNoteEntry"):
        assert name in meth(stub) and name in meth(cuda), name
    fields = lambda s: re.search(r"type GPUOptions struct \{(.*?)\n\}", s, flags=re.S).group(1)
    names = lambda s: re.findall(r"^\s*(\w+)\s+u?int", fields(s), flags=re.M)
    assert names(cuda) == names(stub)
