"""Cross-modal operators on either side of the UNetSCN path, host-side mirrors of the
reference's Python code over the C ABI in include/mopa_xm.h (same names, argument meaning and return values):

  lift_and_classify   Net2DSeg.forward's 2D -> 3D lifting loop + linear heads   mopa/models/xmuda_arch.py:62-77
  xm_kl_div           the cross-modal KL loss of the train scripts              mopa/train/train_xmuda_mopa.py:389-398, 440-445
  mask_cons_loss      SAM mask-consistency loss                                 mopa/common/utils/loss.py:241-283
  post_process        VGI post-processing of the mix-matched scans              mopa/data/mixmatch_ss.py:458-559

No CPU fallback: tensors that are not on a CUDA device raise.
"""
from .lifting import lift_and_classify, xm_kl_div
from .losses import mask_cons_loss
from .vgi import post_process

__all__ = ["lift_and_classify", "xm_kl_div", "mask_cons_loss", "post_process"]
