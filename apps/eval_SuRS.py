#!/usr/bin/env python
"""Inference CLI with the reference's flags (reference apps/eval_SuRS.py): for every image in
``--dataroot`` run the encoder and the B200 reconstruction and write ``<results>/<name>/<subject>_{HR,LR}.obj``.

    python apps/eval_SuRS.py --dataroot data --load_netG_checkpoint_path weights/netG_epoch_12 \
        --residual --resolution 512 --loadSize 1024 --b_min -0.5 -0.5 -0.5 --b_max 0.5 0.5 0.5
    torchrun --nproc-per-node 8 apps/eval_SuRS.py ...      # one image per GPU at a time (BASELINE config 4)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from surs_b200 import _capi
from surs_b200.lib.data import EvalImageFolder
from surs_b200.lib.model import SuRSNet
from surs_b200.lib.options import BaseOptions
from surs_b200.lib.train_util import gen_mesh


def main(argv=None):
    opt = BaseOptions().parse(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(opt.gpu_id)))
    cuda = torch.device("cuda:%d" % local)
    torch.cuda.set_device(cuda)
    data = EvalImageFolder(opt)
    net = SuRSNet(opt, precision={"fp32": _capi.PREC_FP32, "fp16": _capi.PREC_FP16, "fp16x3": _capi.PREC_FP16X3, "fp16r": _capi.PREC_FP16R}[opt.precision],
                  encoder_mode=opt.encoder_mode).to(cuda)
    if opt.load_netG_checkpoint_path is not None:
        net.load_state_dict(torch.load(opt.load_netG_checkpoint_path, map_location=cuda))
    net.eval()
    out_dir = os.path.join(opt.results_path, opt.name)
    os.makedirs(out_dir, exist_ok=True)
    done = []
    with torch.no_grad():
        for i in range(rank, len(data), world):                 # subjects are independent: replicas, no collective
            item = data[i]
            path = os.path.join(out_dir, item["name"] + ".obj")
            gen_mesh(opt, net, cuda, item, path, use_octree=not opt.no_octree)
            done.append(path)
    print("rank %d: %d of %d subjects -> %s" % (rank, len(done), len(data), out_dir))
    return done


if __name__ == "__main__":
    main()
