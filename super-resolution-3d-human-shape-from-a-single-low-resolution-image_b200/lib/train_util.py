"""Drop-in for the mesh driver of the reference's ``lib/train_util.py`` / ``lib/net_util.py``."""
import numpy as np
import torch

from .mesh_util import reconstruction, save_obj_mesh


def make_calib(device):
    """reference lib/train_util.py:63-66: diag(2,-2,2,1) as a [1,4,4] float tensor."""
    m = np.identity(4) * 2
    m[1, 1] = -2
    m[3, 3] = 1
    return torch.Tensor(m).float().unsqueeze(0).to(device=device)


def gen_mesh(opt, net, cuda, data, save_path, use_octree=True):
    """reference lib/train_util.py:53-85: encoder forward, reconstruction, two OBJ files
    (``*_HR.obj`` and ``*_LR.obj``)."""
    image_tensor = data['img_LR'].to(device=cuda)
    _, feature_lr, feature_hr = net.super_res(image_tensor)
    net.filter_hr(feature_hr)
    net.filter_lr(feature_lr)
    calib_tensor = make_calib(cuda)
    verts_hr, faces_hr, _, _, verts_lr, faces_lr, _, _ = reconstruction(
        opt, net, cuda, calib_tensor, opt.resolution, data['b_min'], data['b_max'],
        use_octree=use_octree, num_samples=opt.num_samples)
    save_obj_mesh(save_path[:-4] + '_HR.obj', verts_hr, faces_hr)
    save_obj_mesh(save_path[:-4] + '_LR.obj', verts_lr, faces_lr)
