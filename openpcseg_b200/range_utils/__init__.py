"""``range_utils`` operator surface of the reference's range_lib (RPVNet's point -> range-image
scatter-mean; package/range_lib.zip = pcseg/model/segmentor/fusion/rpvnet/range_lib) on the libb2s
kernels.  ``openpcseg_b200.install_as_torchsparse()`` registers it as ``range_utils`` so that
``import range_utils.nn.functional as rnf`` (rpvnet.py:26) resolves here."""
from . import nn  # noqa: F401

__version__ = "1.0.0"
