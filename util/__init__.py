"""Drop-in `util` package: `from util import metric` / `from util.metric import Metric, ECELoss` (reference
train.py:15-17).  Only `metric` is provided — `util/utils.py` of the reference (checkpoint / wandb / argparse
glue) has no numerics and stays the caller's own file."""
