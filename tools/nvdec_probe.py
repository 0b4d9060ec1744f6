"""Diagnostic: can this box's libnvcuvid see an NVDEC engine?  (driver-level, no torch)"""
import ctypes, os, subprocess, struct
print(subprocess.run("nvidia-smi --query-gpu=name,driver_version --format=csv,noheader; ls -la /usr/lib/libnvcuvid* /usr/local/nvidia/lib/libnvcuvid* /usr/lib/x86_64-linux-gnu/libnvcuvid* 2>&1; echo CAPS=$NVIDIA_DRIVER_CAPABILITIES; ls /dev | grep -i nvidia; ls /dev/nvidia-caps 2>&1 | head", shell=True, capture_output=True, text=True).stdout)
cu = ctypes.CDLL("libcuda.so.1")
print("cuInit", cu.cuInit(0))
dev = ctypes.c_int(0); print("cuDeviceGet", cu.cuDeviceGet(ctypes.byref(dev), 0))
ctx = ctypes.c_void_p(); print("retain", cu.cuDevicePrimaryCtxRetain(ctypes.byref(ctx), dev)); print("setcur", cu.cuCtxSetCurrent(ctx))
for libname in ("libnvcuvid.so.1", "/usr/local/nvidia/lib/libnvcuvid.so.1"):
    try:
        nv = ctypes.CDLL(libname)
    except OSError as e:
        print(libname, "load failed", e); continue
    for codec in (4, 8, 11, 5, 2, 0):
        buf = (ctypes.c_ubyte * 256)()
        struct.pack_into("<iiI", buf, 0, codec, 1, 0)
        rc = nv.cuvidGetDecoderCaps(buf)
        sup, nnv, mask, mw, mh = struct.unpack_from("<BBHII", buf, 24)
        print(libname, "codec", codec, "rc", rc, "supported", sup, "engines", nnv, "max", mw, mh)
print(subprocess.run("nvidia-smi -q | grep -i -B1 -A4 'decoder\\|encoder' | head -40; nvidia-smi -q | grep -i -A3 'virtualization'; cat /proc/driver/nvidia/version 2>&1 | head -3; ls -la /proc/driver/nvidia/capabilities 2>&1 | head; env | grep -i nvidia", shell=True, capture_output=True, text=True).stdout)
# create a decoder directly (CUVIDDECODECREATEINFO, zero tail)
nv = ctypes.CDLL("libnvcuvid.so.1")
ci = (ctypes.c_ubyte * 512)()
struct.pack_into("<QQQiiQQQQQQhhhhiiQQQQ", ci, 0, 176, 112, 8, 4, 1, 4, 0, 0, 176, 112, 0, 0, 0, 176, 112, 0, 0, 176, 112, 2, 0)
h = ctypes.c_void_p()
print("cuvidCreateDecoder rc", nv.cuvidCreateDecoder(ctypes.byref(h), ci), h.value)
lock = ctypes.c_void_p()
print("cuvidCtxLockCreate rc", nv.cuvidCtxLockCreate(ctypes.byref(lock), ctx))
