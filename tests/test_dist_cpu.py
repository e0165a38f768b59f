"""World-size-2 CPU (gloo) tests of the host-side logic of the N>1 path: the ncclUniqueId hand-off through
torch.distributed, the data/seed partition, and the loss/gradient normalisation the fused optimizer applies after
the sum all-reduce (grad_scale = 1/world) versus a plain DDP average."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from llark_b200.comm import broadcast_unique_id, shard_seeds

    uid = broadcast_unique_id(rank)
    seeds = shard_seeds(rank, world, 3, base=10)
    # gradient semantics: sum-all-reduce then scale by 1/world == DDP mean of per-rank mean losses
    g = torch.Generator().manual_seed(seeds[0])
    local_grad = torch.randn(64, generator=g)
    summed = local_grad.clone()
    dist.all_reduce(summed, op=dist.ReduceOp.SUM)
    scaled = summed * (1.0 / world)
    gathered = [torch.zeros(64) for _ in range(world)]
    dist.all_gather(gathered, local_grad)
    q.put((rank, uid, seeds, torch.allclose(scaled, torch.stack(gathered).mean(0), atol=1e-6)))
    dist.destroy_process_group()


def test_world2_unique_id_seeds_and_grad_scale():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, uid0, s0, ok0), (r1, uid1, s1, ok1) = res
    assert len(uid0) == 128 and uid0 == uid1 and any(uid0)  # both ranks hold rank 0's ncclUniqueId
    assert s0 == [10, 12, 14] and s1 == [11, 13, 15]  # disjoint, interleaved
    assert ok0 and ok1


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU-only box and exports every symbol include/llark_b200.h declares."""
    import re

    sys.path.insert(0, ROOT)
    from llark_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "llark_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(llark_[a-z0-9_]+)\s*\(", hdr)))
    lib = _lib.lib()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == declared


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/llark_b200.h against the argtypes / restype llark_b200/_lib.py binds: same arity, and
    each parameter in the same class (pointer / stream / comm handle -> void*, long long, int, float). Catches ABI drift
    of entry points that cannot be called in the CPU suite."""
    import ctypes
    import re

    sys.path.insert(0, ROOT)
    from llark_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "llark_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", " ", hdr)
    protos = re.findall(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(llark_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr)
    assert len(protos) >= 40

    def classify(decl):
        decl = decl.strip()
        if decl in ("", "void"):
            return None
        if "*" in decl or re.search(r"\b(cudaStream_t|ncclComm_t|LlarkComm)\b", decl):
            return ctypes.c_void_p
        if re.search(r"\blong\s+long\b|\bsize_t\b|\bint64_t\b", decl):
            return ctypes.c_longlong
        if re.search(r"\bfloat\b", decl):
            return ctypes.c_float
        if re.search(r"\bdouble\b", decl):
            return ctypes.c_double
        if re.search(r"\b(int|unsigned)\b", decl):
            return ctypes.c_int
        raise AssertionError(f"unclassified parameter {decl!r}")

    lib = _lib.lib()
    seen = set()
    for ret, name, args in protos:
        seen.add(name)
        want = [c for c in (classify(a) for a in args.split(",")) if c is not None]
        fn = getattr(lib, name)
        have = list(fn.argtypes or [])
        assert len(have) == len(want), (name, len(have), len(want))
        for i, (h, w) in enumerate(zip(have, want)):
            assert h is w, (name, i, h, w)
        ret = ret.replace("extern", "").strip()
        want_ret = None if ret == "void" else classify(ret)
        assert fn.restype is want_ret, (name, fn.restype, want_ret)
    assert seen == set(_lib.SYMBOLS)


def test_every_call_site_passes_as_many_arguments_as_the_binding_declares():
    """Static arity check of every `lib().llark_*(...)` call in the package against the bound argtypes: ctypes would only
    complain at call time, which for GPU entry points is never in the CPU suite."""
    import ast
    import glob

    sys.path.insert(0, ROOT)
    from llark_b200 import _lib

    lib = _lib.lib()
    checked = 0
    files = glob.glob(os.path.join(ROOT, "llark_b200", "**", "*.py"), recursive=True)
    files += [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for path in files:
        for node in ast.walk(ast.parse(open(path).read())):
            if not (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute)
                    and node.func.attr.startswith("llark_")):
                continue
            name = node.func.attr
            assert not node.keywords and not any(isinstance(a, ast.Starred) for a in node.args), (path, node.lineno)
            fn = getattr(lib, name)
            assert len(node.args) == len(fn.argtypes or []), (path, node.lineno, name, len(node.args))
            checked += 1
    assert checked >= 45


def test_no_cpu_fallback():
    """Calling an op without CUDA tensors must fail loudly (the product path never routes through the oracle)."""
    import pytest

    sys.path.insert(0, ROOT)
    from llark_b200 import ops

    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    src = open(os.path.join(ROOT, "llark_b200", "engine.py")).read() + open(
        os.path.join(ROOT, "llark_b200", "ops.py")).read()
    assert "oracle" not in src
    # no module of the package imports the oracle, the tests or the CPU stand-in -- not even lazily inside a function
    import ast
    import glob

    for path in glob.glob(os.path.join(ROOT, "llark_b200", "**", "*.py"), recursive=True):
        for node in ast.walk(ast.parse(open(path).read())):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            for n in names:
                root = n.split(".")[0]
                assert root not in ("oracle", "tests", "cpu_ops"), (path, node.lineno, n)


def test_cosine_schedule_matches_hf_formula():
    from llark_b200.trainer import cosine_with_warmup
    import math

    total, warm = 1000, 30
    assert cosine_with_warmup(0, total, warm) == 0.0
    assert abs(cosine_with_warmup(15, total, warm) - 0.5) < 1e-9
    assert abs(cosine_with_warmup(30, total, warm) - 1.0) < 1e-9
    mid = 30 + (1000 - 30) // 2
    assert abs(cosine_with_warmup(mid, total, warm) - 0.5 * (1 + math.cos(math.pi * (mid - 30) / 970))) < 1e-12
    assert cosine_with_warmup(1000, total, warm) < 1e-9
