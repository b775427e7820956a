"""Command-line driver: the flag surface of the reference's src/train_rec.py (:17-46 - same names,
same defaults, same per-regulariser loop and console block), for the models on the hot path
(``--rec bprmf | vbpr | gradfashion``).

    cd src && python -m fvx.train_rec --dataset amazon_men --rec vbpr --gpu 0 ...

Deviations, all documented in DESIGN.md: ``--validation`` parses real booleans (the reference's
``type=bool`` makes any non-empty string true, :29); ``--gpu`` selects the CUDA device (the
reference's default -1 means CPU, which does not exist here, so the default is 0) and ``--rec``
defaults to ``vbpr``; new optional flags ``--adam_mode``, ``--sampler``, ``--seed``,
``--tensor_cores`` (default on), ``--graph_steps`` (default on), ``--data_root``, ``--results_root``.
"""
import argparse

from .config import configs


def _flag(text):
    return str(text).strip().lower() not in ("0", "false", "no", "off", "")


# (name, type, default, extra argparse keywords).  The first block is the reference's surface.
REFERENCE_FLAGS = (
    ("gpu", int, 0, {}),
    ("best_metric", str, "ndcg", {}),
    ("dataset", None, "amazon_baby", {"nargs": "?"}),
    ("rec", None, "vbpr", {"nargs": "?"}),
    ("batch_size", int, 256, {}),
    ("top_k", int, 20, {}),
    ("epochs", int, 200, {}),
    ("verbose", int, -1, {}),
    ("batch_eval", int, 128, {}),
    ("lr", float, 0.001, {}),
    ("validation", _flag, True, {}),
    ("restore_epochs", int, 1, {}),
    ("list_of_regs", float, [0.0], {"nargs": "+"}),
    ("cnn_model", None, "vgg19", {"nargs": "?"}),
    ("output_layer", None, "fc2", {"nargs": "?"}),
    ("embed_k", int, 128, {}),
    ("embed_d", int, 20, {}),
    ("reg", float, 0, {}),
    # GradFashion.py:30-31 reads these two; the reference's own parser never defines them (its --rec gradfashion run
    # fails with AttributeError): added here with the sizes of its experiments' defaults
    ("embed_color", int, 8, {}),
    ("embed_edges", int, 8, {}),
)
ENGINE_FLAGS = (
    ("adam_mode", None, "auto", {"choices": ["auto", "deferred", "dense", "lazy"]}),
    ("sampler", None, "host_ref", {"choices": ["host_ref", "device"]}),
    ("seed", int, 0, {}),
    ("tensor_cores", _flag, True, {}),
    ("graph_steps", _flag, True, {}),
    ("data_root", None, None, {}),
    ("results_root", None, None, {}),
)


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="BPRMF / VBPR training on the fvx engine (reference CLI surface)")
    for name, kind, default, extra in REFERENCE_FLAGS + ENGINE_FLAGS:
        kw = dict(extra, default=default)
        if kind is not None:
            kw["type"] = kind
        ap.add_argument("--" + name, **kw)
    return ap.parse_args(argv)


def _model_class(rec):
    from .recommender.models.BPRMF import BPRMF
    from .recommender.models.GradFashion import GradFashion
    from .recommender.models.VBPR import VBPR
    table = {"bprmf": BPRMF, "vbpr": VBPR, "gradfashion": GradFashion}
    if rec not in table:
        raise NotImplementedError('Not implemented or unknown Recommender Model.')
    return table[rec]


def train(argv=None):
    """One full ``model.train()`` per entry of ``--list_of_regs`` (train_rec.py:60-90); returns the
    list of results dicts."""
    from .dataset import dataset as _dataset
    from .dataset.dataset import DataLoader
    args = parse_args(argv)
    # the host_ref sampler streams are seeded once per run and consumed across the regulariser iterations, as the
    # reference's global streams are (BPRMF.py:15-16 seeds at import; dataset.py:83-114 draws from them)
    _dataset._SHARED_STREAMS.clear()
    args.share_sampler_streams = True
    configs.set_roots(data=args.data_root, results=args.results_root)
    args.device = "cuda:%d" % max(args.gpu, 0)
    regs = list(args.list_of_regs)
    rule = '-' * 68
    runs = []
    for n, reg in enumerate(regs, start=1):
        print(rule)
        print('ITERATION %d/%d WITH REGULARIZATION: %f' % (n, len(regs), reg))
        data = DataLoader(params=args)
        print("Training {0} on {1}".format(args.rec, args.dataset))
        args.reg = reg
        print("Parameters:")
        print("".join("\t- %s = %s\n" % (key, value) for key, value in vars(args).items()))
        runs.append(_model_class(args.rec)(data, args).train())
        print('END REGULARIZATION')
        print(rule)
    return runs


if __name__ == '__main__':
    train()
