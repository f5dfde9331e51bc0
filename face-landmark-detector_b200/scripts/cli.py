"""face_landmark_detect CLI — mirror of reference scripts/cli.py.

`predict` keeps the reference's options (cli.py:68-74) and, unlike the reference (which ignores --net), honours
--net when no <checkpoints_path>_config.json exists (random-init weights, for smoke runs).  `train` is out of
scope of this build and says so."""
import os
import sys
try:
    import keypoints_detector  # noqa: F401
except ModuleNotFoundError:
    sys.path.append(os.path.realpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), os.pardir)))
from keypoints_detector import prediction, training
import click
from time import time
from functools import wraps


def timing(f):
    @wraps(f)
    def _wrap_func(*args, **kw):
        ts = time()
        result = f(*args, **kw)
        te = time()
        print('func:%r args:[%r, %r] took: %2.4f sec' % (f.__name__, args, kw, te - ts))
        return result
    return _wrap_func


@click.command()
@click.option('--data_dir', default="./", type=str, help="Training data location")
@click.option("--checkpoints_path", type=str, default="./weights", help="Keypoints model path")
def train(data_dir, checkpoints_path):
    raise click.ClickException("training is out of scope of the B200 inference build; train with the reference and export "
                               "the weights to .npz (see INTEGRATION.md)")


@click.command()
@click.option("--checkpoints_path", type=str, default=None, help="Keypoints model path")
@click.option('--inp', required=True, type=str, help="Image (or directory of images) to predict")
@click.option('--net', default='default', type=click.Choice(list(training.LANDMARKS_MODELS.keys())), help="Default network")
@click.option('--out_fname', default=None, type=str, help="Where to write the coloured class map")
@timing
def predict(checkpoints_path, inp, net, out_fname):
    if checkpoints_path is not None and os.path.isfile(checkpoints_path + "_config.json"):
        return prediction.keypts_predict(inp=inp, checkpoints_path=checkpoints_path, out_fname=out_fname)
    model = training.LANDMARKS_MODELS[net](68).init_weights(0)
    print("no checkpoint given/found: using random-init %s" % net)
    return prediction.keypts_predict(model=model, inp=inp, out_fname=out_fname)


@click.group()
def main():
    return 0


main.add_command(train, "train")
main.add_command(predict, "predict")

if __name__ == "__main__":
    sys.exit(main())
