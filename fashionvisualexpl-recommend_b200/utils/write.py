"""Result persistence with the reference's file names (src/utils/write.py:14-22)."""
import pickle


def save_obj(obj, name):
    """Pickle ``obj`` to ``name + '.pkl'`` (the results-metrics dict of BPRMF.train)."""
    with open(name + ".pkl", "wb") as f:
        pickle.dump(obj, f)
