"""`python -m cleanba_b200.cleanba_impala ...` -- drop-in for `python cleanba/cleanba_impala.py ...` (V-trace learner,
PyTorch-style RMSProp, concurrency on by default, num_steps 20; cleanba_impala.py:60-87)."""
import tyro

from .cleanba_ppo import main
from .sebulba import Args, impala_defaults

if __name__ == "__main__":
    main(tyro.cli(Args, default=impala_defaults(Args())))
