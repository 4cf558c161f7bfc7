"""Re-export of the package's synthetic-workload generators for the test-suite and golden scripts."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "parallel-genomeseq_b200", "synth.py")
_spec = importlib.util.spec_from_file_location("pgs_synth", _p)
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
globals().update({k: v for k, v in vars(_m).items() if not k.startswith("__")})
