"""Drop-in for the reference's `run_multiple/final_main_iteration_{wb,ca}.py` sweeps (same flags, same CSV).
Single GPU: `python run_multiple.py ...`; one box: `torchrun --nproc-per-node N run_multiple.py ...` (members sharded
across the GPUs).  See debiasing-multi-modal_b200/sweep.py."""
import dbmm
from dbmm.sweep import main

if __name__ == "__main__":
    main()
