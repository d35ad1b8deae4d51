"""Execute the reference-side binding printed in INTEGRATION.md (section B: a `PhysicalSystem` subclass a reference maintainer would add,
forwarding to libgemb200.so through ctypes) against the unmodified reference: build a reference SCMLSystem, hand it to the stub, and
let it fill `gemb200_config` and call `gemb200_create`.  Without a GPU the call must get past validation and stop at the CUDA stage.
Container-only: needs /root/reference."""
import re, sys, warnings
warnings.filterwarnings("ignore"); sys.dont_write_bytecode=True
HERE = __file__.rsplit("/", 2)[0]
sys.path.insert(0, HERE + "/_shims"); sys.path.insert(0, "/root/reference/src"); sys.path.insert(0, HERE.rsplit("/", 1)[0])
md = open(HERE.rsplit("/", 1)[0] + "/INTEGRATION.md").read()
code = re.search(r"```python\n(# gym_electric_motor/physical_systems/b200_system.py.*?)```", md, re.S).group(1)
ns = {}
exec(compile(code, "integration_stub", "exec"), ns)
import gym_electric_motor as gem
from gym_electric_motor.core import ElectricMotorVisualization
class NoViz(ElectricMotorVisualization): pass
env = gem.make("Cont-CC-PMSM-v0", visualization=NoViz())
try:
    sys_ = ns["B200SCMLSystem"](env.physical_system.unwrapped, num_envs=16)
    print("created (GPU present)")
except Exception as e:
    print(type(e).__name__, str(e)[:200])
