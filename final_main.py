"""Drop-in for the reference's `python final_main.py ...` (same flags, same artefacts), running the
B200-native adapter path.  See debiasing-multi-modal_b200/cli.py."""
import dbmm
from dbmm.cli import parse_option, train_all_epochs  # noqa: F401  (re-exported like the reference module)
from dbmm.modules import Adapter, CustomCLIP, LinearClassifier, MultipleAdapter, get_text_embedding  # noqa: F401
from dbmm.engine import (balance_val, train_one_epoch, train_reg_one_epoch, train_reg_seq_one_epoch,  # noqa: F401
                         validate, validate_zs)

if __name__ == "__main__":
    opt = parse_option()
    train_all_epochs(opt)
