"""One launch of the front kernel on identical fronts (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge
pkg = ge.load_package(); abi = pkg.abi
eng = pkg.Engine(0)
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
opts = abi.SetupOpts.default(L_minus_override=L) if L > 0 else None
desc, _ = pkg.setup_front(3, [1, .98, .96, .04, .02], opts)
descs = (abi.FrontDesc * nf)(*[desc] * nf)
out = eng.frame_count_batch(descs)
print("kernel ms", eng.last_kernel_ms(), "steps", out[0].steps, "zeros", out[0].zero_count)
