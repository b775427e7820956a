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

    # GradFashion's two handcrafted descriptors (visual_loader_mixin.py:51-54, 60-69): one global max-abs scale each
    def process_color_visual_features(self):
        feats = getattr(self.data, "color_features_raw", None)     # in-memory (tests, benchmarks)
        if feats is None:
            feats = np.load(configs.hist_color_features_path(self.data.params.dataset))
        self.color_features = (feats / np.max(np.abs(feats))).astype(np.float32)
        self.dim_color_features = self.color_features.shape[1]

    def process_edge_visual_features(self):
        p = self.data.params
        feats = getattr(self.data, "edge_features_raw", None)
        if feats is None:
            feats = np.load(configs.edge_features_path(p.dataset, p.cnn_model, p.output_layer))
        self.edge_features = (feats / np.max(np.abs(feats))).astype(np.float32)
        self.dim_edge_features = self.edge_features.shape[1]
