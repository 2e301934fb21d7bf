"""Command line of the reference script (CFFM.py:24-78 flags, :650-695 main), driving the CUDA
engine.  Every reference flag is accepted with the same name, type and default; new flags are
additive (--precision, --device, --seed, --eval_batch)."""
from __future__ import annotations

import argparse
import ast
import logging
from time import time

# (flag, type, default, help) -- same surface as the reference's parse_args
_FLAGS = [
    ("--path", None, "data/", "Input data path."),
    ("--dataset", None, "frappe", "Choose a dataset."),
    ("--epoch", int, 50, "Number of epochs."),
    ("--pretrain", int, 0, "1: initialize from the saved state; 0: random init; -1: save the state every epoch."),
    ("--batch_size", int, 1024, "Batch size."),
    ("--inner_dims", int, 32, "Number of inner dimensions."),
    ("--outer_dims", int, 32, "Number of outer dimensions."),
    ("--lamda", float, 0, "Regularizer for bilinear part."),
    ("--keep", None, "[1.0,1.0]", "Keep probability per layer (accepted, unused by the graph)."),
    ("--lr", float, 0.05, "Learning rate."),
    ("--loss_type", None, "square_loss", "square_loss, log_loss, mse, mae or hybrid."),
    ("--optimizer", None, "AdagradOptimizer", "AdamOptimizer, AdagradOptimizer, GradientDescentOptimizer, MomentumOptimizer."),
    ("--verbose", int, 1, "Show the results per X epochs."),
    ("--batch_norm", int, 0, "Accepted, unused by the graph."),
    ("--tensorboard", int, 0, "Accepted, ignored (the reference's tensorboard branch does not run)."),
    ("--num_field", int, 3, "Valid dimension of the dataset (frappe=10, ml-tag=3, book-crossing=6)."),
    ("--linear_att", int, 1, "Linear attention part (0 disable or 1 enable)."),
    ("--att_dim", int, 0, "Dimension of linear attention (0: same as num_field)."),
    ("--lamda_att", float, 1.0, "Softmax temperature of the linear attention part."),
    ("--inner_conv", int, 1, "Inner convolution part (0 disable or 1 enable)."),
    ("--gamma_inner", int, 1.0, "Accepted, unused by the graph."),
    ("--outer_conv", int, 1, "Outer convolution part (0 disable or 1 enable)."),
    ("--beta_outer", int, 1.0, "Weight of the outer convolution component."),
    ("--activation", None, "relu", "relu, prelu, elu, selu or gelu."),
]
_EXTRA = [
    ("--precision", None, "fp32", "Arithmetic of the conv contraction: fp32 or bf16."),
    ("--device", int, 0, "CUDA device ordinal."),
    ("--seed", int, 2021, "Seed of the initialisers and of the batch sampler (reference: unseeded)."),
    ("--eval_batch", int, 0, "Batch size used by evaluate() (0: same as --batch_size, like the reference)."),
]


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Run CFFM.")
    for flag, typ, default, hlp in _FLAGS + _EXTRA:
        if typ is None:
            parser.add_argument(flag, nargs="?", default=default, help=hlp)
        else:
            parser.add_argument(flag, type=typ, default=default, help=hlp)
    return parser.parse_args(argv)


def configure_logging(log_filename):
    """CFFM.py:81-94: file 'logging.log' (append, DEBUG) + console (INFO), same line format."""
    logging.basicConfig(level=logging.DEBUG, format="%(asctime)s %(filename)s:%(message)s",
                        datefmt="%Y-%m-%d %A %H:%M:%S", filename=log_filename, filemode="a")
    console = logging.StreamHandler()
    console.setLevel(logging.INFO)
    console.setFormatter(logging.Formatter("%(asctime)s %(filename)s:%(message)s"))
    logging.getLogger().addHandler(console)


def main(argv=None):
    from .data import LoadData
    from .model import CFFM

    args = parse_args(argv)
    configure_logging("logging.log")
    keep = ast.literal_eval(args.keep)
    if args.verbose > 0:
        logging.info(
            "CFFM: dataset=%s, factors=%d, loss_type=%s, #epoch=%d, batch=%d, lr=%.4f, lambda=%.1e, keep=%s, optimizer=%s"
            ", batch_norm=%d, num_field=%d, linear_att=%d, att_dim=%d,lamda_att=%.2f,inner_conv=%d,gamma_inner=%.1f,outer_conv=%d,"
            "beta_outer=%.1f, activation=%s"
            % (args.dataset, args.inner_dims, args.loss_type, args.epoch, args.batch_size, args.lr, args.lamda, keep,
               args.optimizer, args.batch_norm, args.num_field, args.linear_att, args.att_dim, args.lamda_att,
               args.inner_conv, args.gamma_inner, args.outer_conv, args.beta_outer, args.activation))
    data = LoadData(args.path, args.dataset, args.loss_type)
    save_file = "pretrain/CFFM/%s_%d/%s_%d" % (args.dataset, args.inner_dims, args.dataset, args.inner_dims)
    t1 = time()
    model = CFFM(data.features_M, args.pretrain, save_file, args.inner_dims, args.outer_dims, args.loss_type, args.epoch,
                 args.batch_size, args.lr, args.lamda, keep, args.optimizer, args.batch_norm, args.verbose,
                 args.tensorboard, args.num_field, args.linear_att, args.att_dim, args.lamda_att, args.inner_conv,
                 args.gamma_inner, args.outer_conv, args.beta_outer, args.activation, random_seed=args.seed,
                 precision=args.precision, device=args.device, eval_batch=args.eval_batch or None, batch_seed=args.seed)
    model.train(data)
    # CFFM.py:681-695 -- best epoch by validation RMSE, then by validation R2
    best_valid_score = min(model.valid_rmse)
    best_epoch = model.valid_rmse.index(best_valid_score)
    logging.info("Best Iter of RMSE (validation)= %d train = %.4f, valid = %.4f, test = %.4f [%.1f s]"
                 % (best_epoch + 1, model.train_rmse[best_epoch], model.valid_rmse[best_epoch],
                    model.test_rmse[best_epoch], time() - t1))
    best_r2 = model.valid_r2.index(max(model.valid_r2))
    logging.info("Best Iter of R2 (validation)= %d train = %.4f, valid = %.4f, test = %.4f [%.1f s]"
                 % (best_epoch + 1, model.train_r2[best_r2], model.valid_r2[best_r2], model.test_r2[best_r2],
                    time() - t1))
    return model


if __name__ == "__main__":
    main()
