"""Driver used by tests/test_gpu_reference_models.py through `python -m lss2_multimodal_nu_b200.run`:
builds the reference's PreTrainingModel (pre_train_vovnet.py:29-173, a class the class-level patch
cannot see) and runs one forward; writes the output and whether the instance got patched.

    python -m lss2_multimodal_nu_b200.run tests/scripts/run_pretraining_model.py OUT.pt [--no-run-hook]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import ref_import  # noqa: E402
from lss2_multimodal_nu_b200 import synthetic as S  # noqa: E402


def main(out_path):
    ptv = ref_import.load_script("pre_train_vovnet")
    cfg = S.config("config1", B=2)
    torch.manual_seed(7)
    model = ptv.PreTrainingModel(cfg.B, cfg.grid_conf(), cfg.data_aug_conf(), outC=4, vovnet_type="vovnet39",
                                 pretrained=False, lss_version="v1").cuda().eval()
    cal = {k: torch.from_numpy(v).cuda() for k, v in S.make_calibration(cfg, 5).items()}
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    imgs = torch.randn(cfg.B, cfg.N, 3, *cfg.final_dim, device="cuda", generator=g)
    with torch.no_grad():
        out = model(imgs, cal["rots"], cal["trans"], cal["intrins"], cal["post_rots"], cal["post_trans"])
    torch.save({"out": out.cpu(), "patched": bool(model.__dict__.get("_lss_b200_installed", False)),
                "state": {k: v.cpu() for k, v in model.state_dict().items()}}, out_path)


if __name__ == "__main__":
    main(sys.argv[1])
