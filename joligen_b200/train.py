"""joliGEN's own `train.py`, unmodified, with the B200 kernels under it:

    cd /path/to/joliGEN && python -m joligen_b200.train --config_json examples/example_ddpm_mario.json ...
    (or JOLIGEN_ROOT=/path/to/joliGEN python -m joligen_b200.train ...)

This is the binding INTEGRATION.md section 2 describes, installed from the OUTSIDE so that the reference tree stays
untouched: `models.create_model` (train.py:195) is wrapped, and every network of the created model that `accelerate()`
knows (diffusion generators over UNet / UNetVid / UNetGeneratorRefAttn, the b2b generator, ResnetGenerator,
NLayerDiscriminator) is swapped for its B200 mirror right after construction — before `model.setup()`, before
`parallelize()` wraps the nets in DistributedDataParallel, with the optimizers the model already built still holding the
very same Parameter objects.  Everything else (options, data loading, visualizer, checkpoints, metrics) is joliGEN's.

One process per GPU as in the reference: `launch_training` spawns `train_gpu` (train.py:518-545); the spawned
interpreters import THIS module's `train_gpu`, which installs the hook there before handing over.
"""
import argparse
import os
import sys

_ORIG_TRAIN_GPU = None


def accelerate_model(model, verbose=True):
    """Swap every net of a joliGEN model (BaseModel.model_names -> `net<name>`) that has a B200 mirror.  Returns the
    names that were swapped."""
    from .accelerate import accelerate
    swapped = []
    for name in list(getattr(model, "model_names", [])):
        if not isinstance(name, str):
            continue
        net = getattr(model, "net" + name, None)
        if net is None:
            continue
        new = accelerate(net)
        if new is not net:
            setattr(model, "net" + name, new)
            swapped.append(name)
    if verbose:
        print("joligen_b200: networks on the B200 path: %s" % (", ".join(swapped) if swapped else "none"))
    return swapped


def install_hook():
    """Wrap create_model where train.py and the models package look it up (idempotent)."""
    import models
    import train as ref_train
    if getattr(ref_train, "_jg_hooked", False):
        return
    orig = ref_train.create_model

    def create_model(opt, rank):
        model = orig(opt, rank)
        accelerate_model(model, verbose=(rank == 0))
        return model

    ref_train.create_model = create_model
    models.create_model = create_model
    ref_train._jg_hooked = True


def train_gpu(rank, world_size, opt, trainset, trainset_temporal):
    """The spawn target (picklable by module path): hook, then the reference's train_gpu."""
    import train as ref_train
    install_hook()
    target = _ORIG_TRAIN_GPU if _ORIG_TRAIN_GPU is not None else ref_train.train_gpu
    return target(rank, world_size, opt, trainset, trainset_temporal)


def _reference_on_path():
    root = os.environ.get("JOLIGEN_ROOT", os.getcwd())
    if not os.path.exists(os.path.join(root, "train.py")) or not os.path.isdir(os.path.join(root, "models")):
        raise SystemExit("joligen_b200.train: run from a joliGEN checkout or set JOLIGEN_ROOT (no train.py / models in %s)"
                         % root)
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


def main(argv=None):
    global _ORIG_TRAIN_GPU
    _reference_on_path()
    import train as ref_train
    install_hook()
    if _ORIG_TRAIN_GPU is None:
        _ORIG_TRAIN_GPU = ref_train.train_gpu
        ref_train.train_gpu = train_gpu    # what launch_training hands to mp.spawn (train.py:537-543)
    parser = argparse.ArgumentParser(add_help=False)   # train.py:556-566
    parser.add_argument("--config_json", type=str, default="", help="path to json config")
    main_opt, remaining = parser.parse_known_args(argv)
    opt = ref_train.get_opt(main_opt, remaining)
    ref_train.launch_training(opt)


if __name__ == "__main__":
    main()
