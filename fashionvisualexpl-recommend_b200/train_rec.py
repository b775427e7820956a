"""Command-line driver with the flags and defaults of the reference's src/train_rec.py
(:17-46), restricted to the models on the hot path (``--rec bprmf | vbpr``).

    cd src && python -m fvx.train_rec --dataset amazon_men --rec vbpr --gpu 0 ...

Deviations, all documented in DESIGN.md: ``--validation`` parses real booleans (the
reference's ``type=bool`` makes any non-empty string true, :29); ``--gpu`` selects the
CUDA device (the reference's default -1 means CPU, which does not exist here);
new optional flags ``--adam_mode``, ``--sampler``, ``--seed``, ``--tensor_cores``,
``--data_root``, ``--results_root``.
"""
import argparse

from .config import configs


def _bool(s):
    return str(s).lower() not in ("0", "false", "no", "")


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Run train of the Recommender Model.")
    parser.add_argument('--gpu', type=int, default=0)
    parser.add_argument('--best_metric', type=str, default='ndcg')
    parser.add_argument('--dataset', nargs='?', default='amazon_baby', help='dataset name')
    parser.add_argument('--rec', nargs='?', default="vbpr", help="set recommendation model")
    parser.add_argument('--batch_size', type=int, default=256, help='batch_size')
    parser.add_argument('--top_k', type=int, default=20, help='top-k of recommendation.')
    parser.add_argument('--epochs', type=int, default=200, help='Number of epochs.')
    parser.add_argument('--verbose', type=int, default=-1, help='number of epochs to store model parameters.')
    parser.add_argument('--batch_eval', type=int, default=128, help='batch size on items for evaluation.')
    parser.add_argument('--lr', type=float, default=0.001, help='Learning rate.')
    parser.add_argument('--validation', type=_bool, default=True, help='use the validation set')
    parser.add_argument('--restore_epochs', type=int, default=1)
    parser.add_argument('--list_of_regs', nargs='+', type=float, default=[0.0], help='list of regularization terms')
    parser.add_argument('--cnn_model', nargs='?', default='vgg19', help='Model used for feature extraction.')
    parser.add_argument('--output_layer', nargs='?', default='fc2', help='Output layer for feature extraction.')
    parser.add_argument('--embed_k', type=int, default=128, help='Embedding size.')
    parser.add_argument('--embed_d', type=int, default=20, help='size of low dimensionality for visual features')
    parser.add_argument('--reg', type=float, default=0, help='regularization')
    # additions
    parser.add_argument('--adam_mode', default='deferred', choices=['deferred', 'dense', 'lazy'])
    parser.add_argument('--sampler', default='host_ref', choices=['host_ref', 'device'])
    parser.add_argument('--seed', type=int, default=0)
    parser.add_argument('--tensor_cores', type=_bool, default=False)
    parser.add_argument('--data_root', default=None)
    parser.add_argument('--results_root', default=None)
    return parser.parse_args(argv)


def train(argv=None):
    args = parse_args(argv)
    configs.set_roots(data=args.data_root, results=args.results_root)
    args.device = "cuda:%d" % max(args.gpu, 0)
    from .dataset.dataset import DataLoader
    from .recommender.models.BPRMF import BPRMF
    from .recommender.models.VBPR import VBPR
    out = []
    for it, current_reg in enumerate(list(args.list_of_regs)):
        print('--------------------------------------------------------------------')
        print('ITERATION %d/%d WITH REGULARIZATION: %f' % (it + 1, len(list(args.list_of_regs)), current_reg))
        data = DataLoader(params=args)
        print("Training {0} on {1}".format(args.rec, args.dataset))
        print("Parameters:")
        args.reg = current_reg
        for arg in vars(args):
            print("\t- " + str(arg) + " = " + str(getattr(args, arg)))
        print("\n")
        if args.rec == 'bprmf':
            model = BPRMF(data, args)
        elif args.rec == 'vbpr':
            model = VBPR(data, args)
        else:
            raise NotImplementedError('Not implemented or unknown Recommender Model.')
        out.append(model.train())
        print('END REGULARIZATION')
        print('--------------------------------------------------------------------')
    return out


if __name__ == '__main__':
    train()
