"""Byte-flip / truncation fuzz of the native torch.save reader (csrc/ptfile.cpp) under ASAN + UBSAN.

    g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer -pthread \
        -o /tmp/libptfile_asan.so torbi_b200/csrc/ptfile.cpp
    LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tests/fuzz_native_reader.py

Not collected by pytest (CPU only, ~1 minute).  Accepted files are also read back through
torbi_b200_pt_read into an exactly sized buffer, so an out-of-bounds payload would trip ASAN."""
import ctypes, random, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torbi_b200 import _ffi
lib=ctypes.CDLL('/tmp/libptfile_asan.so')
lib.torbi_b200_pt_inspect.restype=ctypes.c_int
lib.torbi_b200_pt_inspect.argtypes=[ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int]
lib.torbi_b200_pt_read.restype=ctypes.c_int
lib.torbi_b200_pt_read.argtypes=[ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
random.seed(7)
os.chdir('/dev/shm')
torch.save(torch.arange(600, dtype=torch.float32).reshape(20,30), 'fa_good.pt')
raw=bytearray(open('fa_good.pt','rb').read())
paths=(ctypes.c_char_p*1)(b'/dev/shm/fa.pt')
ok=0
for it in range(6000):
    b=bytearray(raw); mode=it%4
    if mode==0:
        for _ in range(random.randint(1,4)): b[random.randrange(len(b))]=random.randrange(256)
    elif mode==1: b=b[:random.randrange(len(b))]
    elif mode==2:
        for _ in range(random.randint(1,6)): b[random.randrange(0,min(700,len(b)))]=random.randrange(256)
    else:
        for _ in range(random.randint(1,6)): b[random.randrange(len(b)-400,len(b))]=random.randrange(256)
    open('fa.pt','wb').write(b)
    info=(_ffi.PtInfo*1)()
    lib.torbi_b200_pt_inspect(paths,1,info,1)
    if info[0].status==0 and 0<=info[0].nbytes<=1<<20:
        ok+=1
        buf=ctypes.create_string_buffer(int(info[0].nbytes)+64)
        dst=(ctypes.c_void_p*1)(ctypes.addressof(buf)); cap=(ctypes.c_int64*1)(info[0].nbytes+64)
        lib.torbi_b200_pt_read(paths,info,1,dst,cap,2)
print('asan fuzz done, accepted',ok)
