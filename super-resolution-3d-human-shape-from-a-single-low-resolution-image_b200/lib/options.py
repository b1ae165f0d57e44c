"""The subset of the reference's ``lib/options.py`` (BaseOptions) that inference reads, with the same
flag names and defaults (reference lib/options.py:9-185), so command lines written for
``apps/eval_SuRS.py`` keep working."""
import argparse


class BaseOptions:
    def initialize(self, parser):
        g = parser.add_argument_group("Data")
        g.add_argument("--dataroot", type=str, default="./data", help="folder with image_final/ and mask_final/")
        g.add_argument("--loadSize", type=int, default=512, help="2 x the side of the low-resolution input (README.md:38)")
        g = parser.add_argument_group("Experiment")
        g.add_argument("--name", type=str, default="example")
        g.add_argument("--num_views", type=int, default=1)
        g.add_argument("--gpu_id", type=int, default=0)
        g.add_argument("--resolution", type=int, default=512, help="# of grid nodes per axis in mesh reconstruction")
        g.add_argument("--z_size", type=float, default=200.0, help="z normalization factor")
        g = parser.add_argument_group("Model")
        g.add_argument("--norm", type=str, default="group")
        g.add_argument("--hg_depth", type=int, default=2)
        g.add_argument("--hg_dim", type=int, default=256)
        g.add_argument("--num_stack_lr", type=int, default=3)
        g.add_argument("--num_stack_hr", type=int, default=1)
        g.add_argument("--mlp_dim_lr", nargs="+", default=[321, 1024, 512, 256, 128, 1], type=int)
        g.add_argument("--mlp_dim_hr", nargs="+", default=[322, 1024, 512, 256, 128, 1], type=int)
        g.add_argument("--mlp_res_layers_lr", nargs="+", default=[2, 3, 4], type=int)
        g.add_argument("--mlp_res_layers_hr", nargs="+", default=[2, 3, 4], type=int)
        g.add_argument("--no_residual", action="store_true", help="no skip connection in mlp")
        g.add_argument("--residual", action="store_true", help="apply residual blocks in the super-resolution branch")
        g.add_argument("--n_block", type=int, nargs="+", default=[2, 2, 2])
        g.add_argument("--scale", type=int, default=2)
        g.add_argument("--rgb_range", type=int, default=255)
        g.add_argument("--b_min", nargs="+", type=float, default=[-128.0, -28.0, -128.0])
        g.add_argument("--b_max", nargs="+", type=float, default=[128.0, 228.0, 128.0])
        g.add_argument("--num_samples", type=int, default=50000)
        g.add_argument("--threshold", type=float, default=0.05, help="octree: corner range below which a cell is uniform")
        g.add_argument("--results_path", type=str, default="./results")
        g.add_argument("--load_netG_checkpoint_path", type=str, default=None)
        g.add_argument("--no_octree", action="store_true", help="dense grid instead of the octree (gen_mesh default: octree)")
        g.add_argument("--precision", type=str, default="fp16r", choices=["fp16", "fp16x3", "fp16r", "fp32"],
                       help="fp16r (default): one tensor-core pass everywhere + split hi/lo operands on every node the 0.5 "
                            "iso-surface can depend on (mesh identical to fp16x3; octree / point queries run fp16x3); "
                            "fp16x3: split operands everywhere (three passes, < 1e-4 from the reference); "
                            "fp32: CUDA-core exact mode; fp16: ONE pass -- explicit opt-in, up to 1.5e-2 from the reference")
        g.add_argument("--encoder_mode", type=str, default="fast", choices=["eager", "fast", "bf16"],
                       help="image encoder: fast = TF32 convolutions (PyTorch's GPU default, as the reference) + one CUDA graph "
                            "per input shape; eager = plain module calls; bf16 = channels_last + bf16 autocast (slower, less accurate)")
        return parser

    def parse(self, argv=None):
        return self.initialize(argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)).parse_args(argv)
