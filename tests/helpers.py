"""Shared test helpers (fixture readers; no product code)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TINY = os.path.join(GOLDEN, "tiny")


def read_lists(path, num_users):
    out = [[] for _ in range(num_users)]
    with open(path) as f:
        for line in f:
            a = line.split("\t")
            out[int(a[0])].append(int(a[1]))
    return out


def tiny_dataset():
    with open(os.path.join(TINY, "stats_after_downloading")) as f:
        lines = f.readlines()
    U, I = int(lines[2].split(": ")[1]), int(lines[3].split(": ")[1])
    tr = read_lists(os.path.join(TINY, "trainingset.tsv"), U)
    va = read_lists(os.path.join(TINY, "validationset.tsv"), U)
    te = read_lists(os.path.join(TINY, "testset.tsv"), U)
    F = np.load(os.path.join(TINY, "original", "cnn_features_resnet50_avg_pool.npy"))
    return U, I, tr, va, te, F


def csr(lists):
    row_ptr = np.zeros(len(lists) + 1, dtype=np.int64)
    row_ptr[1:] = np.cumsum([len(x) for x in lists])
    col_file = np.array([i for x in lists for i in x], dtype=np.int64)
    col_sorted = np.array([i for x in lists for i in sorted(x)], dtype=np.int64)
    return row_ptr, col_file, col_sorted


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
