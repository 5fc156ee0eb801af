"""utils/pytorch3d_chamfer_distance.py of the reference (contact score, test_egohmr.py:496-506) on the library's 1-NN
kernel instead of pytorch3d's `knn_points` (a third-party CUDA extension the reference imports but does not vendor).

`chamfer_distance(x, y)` keeps the reference's modified return convention (:214-222): the UNREDUCED squared
nearest-neighbour distances `(cham_x [N,P1], cham_y [N,P2], None)`.  Extra, optional: `y_index` (int [N]) lets all
samples of an image share one scene cloud, so the driver's `.repeat(1, num_samples, 1, 1)` copy is not needed."""
import torch

from .geometry import _engine_for


def chamfer_distance(x, y, x_lengths=None, y_lengths=None, x_normals=None, y_normals=None, weights=None,
                     batch_reduction=None, point_reduction="mean", y_index=None, compute_y=True):
    if x_lengths is not None or y_lengths is not None or x_normals is not None or y_normals is not None or weights is not None:
        raise NotImplementedError("heterogeneous clouds / normals / weights are not used by the reference's driver")
    if batch_reduction is not None:
        raise NotImplementedError("the reference's modified chamfer_distance is only called unreduced (test_egohmr.py:501)")
    if not (x.is_cuda and y.is_cuda):
        raise RuntimeError("chamfer_distance runs on CUDA tensors only (no CPU fallback)")
    x = x.float().contiguous()
    y = y.float().contiguous()
    N = x.shape[0]
    if y_index is None and y.shape[0] != N:
        raise ValueError("y does not have the correct shape.")
    eng = _engine_for(x.device)
    yi = None if y_index is None else y_index.to(device=x.device, dtype=torch.int32).contiguous()
    cham_x = eng.nn_dist_sq(x, y, None, yi, N)
    cham_y = eng.nn_dist_sq(y, x, yi, None, N) if compute_y else None
    return cham_x, cham_y, None
