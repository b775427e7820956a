"""CNN feature loading (src/dataset/visual_loader_mixin.py:22-31): ``np.load`` of the
``[I, D]`` matrix and ONE global scale by ``max(abs(F))``; the cast to fp32 happens when
the model uploads it (VBPR.py:49-51)."""
import numpy as np

from ..config import configs


class VisualLoader:
    def process_cnn_visual_features(self):
        p = self.data.params
        feats = getattr(self.data, "cnn_features_raw", None)       # in-memory (benchmarks)
        if feats is None:
            feats = np.load(configs.cnn_features_path(p.dataset, p.cnn_model, p.output_layer))
        self.cnn_features = (feats / np.max(np.abs(feats))).astype(np.float32)
        self.dim_cnn_features = self.cnn_features.shape[1]
