"""Deformable convolution / deformable PS-ROI pooling wrappers (reference layers/dcn/)."""
