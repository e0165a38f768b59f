"""ORACLE package -- test infrastructure only. CPU restatements of the reference's algorithms (llama_oracle, mpt_oracle,
jukebox_oracle), the shims that import the reference's own classes in place (ref_shims) and the generators of
tests/golden/*.pt (make_golden*). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import it; nothing under llark_b200/ does (tests/test_dist_cpu.py::test_no_cpu_fallback checks that)."""
