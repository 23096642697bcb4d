"""Mirror of ``gymwipe/networking/attenuation_models.py``: free-space path loss (kernel K1)."""
import torch

from gymwipe_b200 import _native as N


class FsplAttenuation:
    """
    ``attenuation_models.py:19-39``: ``20 log10(d) + 20 log10(f) - 147.55`` dB.  As a class it
    marks a band's attenuation model; :meth:`attenuation` evaluates kernel K1 on tensors.
    """

    @staticmethod
    def attenuation(ax, ay, bx, by, frequency):
        """Attenuation in dB between positions ``(ax, ay)`` and ``(bx, by)`` (CUDA float64 tensors)."""
        assert ax.is_cuda and ax.dtype == torch.float64
        out = torch.empty_like(ax)
        stream = torch.cuda.current_stream(ax.device).cuda_stream
        with torch.cuda.device(ax.device):
            N.check(N.lib().gw_fspl_attenuation(ax.contiguous().data_ptr(), ay.contiguous().data_ptr(),
                                                bx.contiguous().data_ptr(), by.contiguous().data_ptr(),
                                                float(frequency), out.data_ptr(), ax.numel(), stream))
        return out
