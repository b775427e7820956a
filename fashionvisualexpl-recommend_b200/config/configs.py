"""On-disk layout of the hot path (the subset of src/config/configs.py:2-33 that
BPRMF / VBPR / Evaluator / DataLoader touch).  Paths are relative to the working
directory ``src/`` exactly as in the reference; ``set_roots`` re-bases them (tests,
benchmarks) without changing the layout below the roots."""

_ROOTS = {"data": "../data", "results": "../results"}


def set_roots(data=None, results=None):
    if data is not None:
        _ROOTS["data"] = data
    if results is not None:
        _ROOTS["results"] = results


def _data(dataset, *tail):
    return "/".join((_ROOTS["data"], dataset) + tail)


def training_path(dataset):        # configs.py:9
    return _data(dataset, "trainingset.tsv")


def validation_path(dataset):      # configs.py:10
    return _data(dataset, "validationset.tsv")


def test_path(dataset):            # configs.py:11
    return _data(dataset, "testset.tsv")


def dataset_info(dataset):         # configs.py:14
    return _data(dataset, "stats_after_downloading")


def cnn_features_path(dataset, cnn_model, output_layer):   # configs.py:17
    return _data(dataset, "original", "cnn_features_%s_%s.npy" % (cnn_model, output_layer))


def edge_features_path(dataset, cnn_model, output_layer):  # configs.py:20
    return _data(dataset, "original", "edge_features_%s_%s.npy" % (cnn_model, output_layer))


def hist_color_features_path(dataset):                     # configs.py:24
    return _data(dataset, "original", "features", "histograms.npy")


def weight_dir():                  # configs.py:32
    return _ROOTS["results"] + "/rec_model_weights"


def results_dir():                 # configs.py:33
    return _ROOTS["results"] + "/rec_results"
