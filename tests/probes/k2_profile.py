"""One short run of the cold pipeline for ncu captures of the K2 sweep (mode from argv: 0/1/2; LEDs; batch)."""
import sys
sys.path.insert(0, ".")
from tests.probes.k2_modes_probe import run
if __name__ == "__main__":
    mode = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n_leds = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
    run(n_leds, B, modes=(mode,))
