"""Re-export of the shared synthetic generators (the package owns them; the oracle may import the
package, never the other way round)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from lcrec_b200.synth import lowrank_map, seeded_weights, synth_items  # noqa: F401,E402
