import sys, os, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mopa_b200 import synth, _lib
from mopa_b200.unet_scn import UNetSCN
from mopa_b200.scn import compiler, functional as F
net = UNetSCN(1).cuda()
prog = compiler.compiled_for(net.sparseModel)
L = prog._lib
for bs in (1, 8):
    c, f = synth.make_batch(bs, 'nuscenes', 0)
    cd = torch.from_numpy(c).cuda(); ch = torch.from_numpy(c).pin_memory(); fd = torch.from_numpy(f).cuda()
    for coords in (ch, cd):
        for it in range(4):
            torch.cuda.synchronize()
            t0 = time.time()
            handle = prog.ensure_handle(0)
            meta = F.Metadata(3, fd.device)
            n_active = (ctypes.c_int64 * prog.n_levels)(); sizes = (ctypes.c_uint64 * 3)()
            _lib.check(L.mopa_scn_Program_prepare(handle, meta._h, coords.data_ptr(), coords.shape[0], 4, 1 if coords.is_cuda else 0, 1, F._stream(), n_active, sizes))
            t1 = time.time()
            act = torch.empty(sizes[0], dtype=torch.uint8, device='cuda'); scratch = torch.empty(sizes[2], dtype=torch.uint8, device='cuda')
            out = torch.empty(coords.shape[0], 16, device='cuda')
            tensors = prog.tensors()
            params = (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
            t2 = time.time()
            _lib.check(L.mopa_scn_Program_forward(handle, meta._h, fd.data_ptr(), 1, params, 1, 1, act.data_ptr(), scratch.data_ptr(), out.data_ptr(), 16, F._stream()))
            t3 = time.time()
            torch.cuda.synchronize(); t4 = time.time()
            grad_arena = torch.empty(sizes[1], dtype=torch.uint8, device='cuda')
            flat = torch.empty(sum(t.numel() for t in tensors), device='cuda')
            ptrs = []; off = 0
            for t, (m, a) in zip(tensors, prog.slots):
                ptrs.append(None if a.startswith('running_') else flat[off:off+t.numel()].data_ptr()); off += t.numel()
            pg = (ctypes.c_void_p * len(tensors))(*ptrs)
            t5 = time.time()
            _lib.check(L.mopa_scn_Program_backward(handle, meta._h, params, pg, 1, 1, act.data_ptr(), grad_arena.data_ptr(), scratch.data_ptr(), out.data_ptr(), 16, None, 1, F._stream()))
            t6 = time.time()
            torch.cuda.synchronize(); t7 = time.time()
        print('batch %d coords %s: prepare %.2f ms | alloc %.2f | fwd submit %.2f, fwd gpu-drain %.2f | bwd prep %.2f | bwd submit %.2f, drain %.2f | sizes %s' % (
            bs, 'dev' if coords.is_cuda else 'host', (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, (t6-t5)*1e3, (t7-t6)*1e3, [int(x)>>20 for x in sizes]))
