"""ctypes binding of libvpd_b200.so (the C ABI in include/vpd_b200.h).

The prototypes are parsed from the header so the binding cannot drift from the
ABI. There is no fallback of any kind: if the shared library is missing or a
call fails, an exception is raised.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(HERE, '..', 'include', 'vpd_b200.h')
LIB_PATH = os.path.join(HERE, 'libvpd_b200.so')

_SCALARS = {
    'int': ctypes.c_int, 'int32_t': ctypes.c_int32, 'int64_t': ctypes.c_int64,
    'uint64_t': ctypes.c_uint64, 'size_t': ctypes.c_size_t, 'float': ctypes.c_float,
    'double': ctypes.c_double, 'long long': ctypes.c_longlong,
}


def parse_header(path=HEADER):
    """-> {name: (restype, [(argname, ctype)])} for every VPD_API prototype."""
    with open(path) as fp:
        text = fp.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    protos = {}
    for m in re.finditer(r'VPD_API\s+([\w\s\*]+?)\s*\b(vpd_\w+)\s*\(([^;]*?)\)\s*;', text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if '*' in ret:
            restype = ctypes.c_char_p if 'char' in ret else ctypes.c_void_p
        elif ret == 'void':
            restype = None
        else:
            restype = _SCALARS[ret]
        argl = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                if '*' in a:
                    argl.append((a.split('*')[-1].strip(), ctypes.c_void_p))
                else:
                    ty, an = a.rsplit(' ', 1)
                    ty = ty.replace('const ', '').strip()
                    # unknown names are opaque handles / function-pointer typedefs
                    argl.append((an, _SCALARS.get(ty, ctypes.c_void_p)))
        protos[name] = (restype, argl)
    return protos


# int-returning entry points whose result is a value, not a status code
_VALUE_INT = {'vpd_abi_version', 'vpd_net_num_bn', 'vpd_net_num_tensors', 'vpd_assemble_tables'}


class VpdError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        if not os.path.exists(LIB_PATH):
            raise VpdError(
                'vpd_b200: {} not found - build it with `python -m vpd_b200.build` '
                '(there is no CPU or PyTorch fallback)'.format(LIB_PATH))
        self._dll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, args) in self.protos.items():
            fn = getattr(self._dll, name)
            fn.restype = restype
            fn.argtypes = [t for _, t in args]
        self.launches = 0          # kernel-launching calls made (bench bookkeeping)

    def last_error(self):
        return self._dll.vpd_last_error().decode()

    def call(self, name, *args):
        """Call a status-returning entry point; raise on failure."""
        fn = getattr(self._dll, name)
        conv = []
        for a, (_, t) in zip(args, self.protos[name][1]):
            if t is ctypes.c_void_p:
                if a is None:
                    conv.append(None)
                elif hasattr(a, 'data_ptr'):
                    conv.append(a.data_ptr())
                elif isinstance(a, (bytes, int)):
                    conv.append(a)
                else:
                    conv.append(a)          # ctypes byref()/pointer/array
            else:
                conv.append(a)
        if len(conv) != len(self.protos[name][1]):
            raise TypeError('{} expects {} arguments, got {}'.format(
                name, len(self.protos[name][1]), len(conv)))
        rc = fn(*conv)
        if self.protos[name][0] is ctypes.c_int and name not in _VALUE_INT and rc != 0:
            raise VpdError('{} failed: {}'.format(name, self.last_error()))
        return rc

    def raw(self, name):
        return getattr(self._dll, name)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib


def stream_ptr(device=None):
    import torch
    return torch.cuda.current_stream(device).cuda_stream


# ---- vpd_stat_acc buffers (include/vpd_b200.h): int64 [..., 2] = (hi, lo), value = hi/16 + lo/2^52
def acc_zeros(shape, device):
    """A zeroed statistics-accumulator buffer of `shape` elements (torch int64 [*shape, 2])."""
    import torch
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    return torch.zeros(shape + (2,), device=device, dtype=torch.int64)


def acc_from_f64(t):
    """float64 tensor -> accumulator buffer holding the same values (to 2^-52 absolute)."""
    import torch
    t = t.double()
    hi = torch.round(t * 16.0)
    lo = torch.round((t - hi / 16.0) * float(1 << 52))
    return torch.stack([hi.to(torch.int64), lo.to(torch.int64)], dim=-1).contiguous()


def acc_to_f64(a):
    """accumulator buffer [..., 2] -> float64 values"""
    return a[..., 0].double() / 16.0 + a[..., 1].double() / float(1 << 52)
