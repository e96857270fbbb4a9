import sys; sys.path.insert(0,"depth-map-fusion-utils_b200")
import dmf_b200 as D, numpy as np, ctypes as C, time
for name in ("S512","S512","S512","S1024"):
    sc=D.scenes.scene(name); ctx=D.Context(0); vol=D.VoxelVolume(ctx)
    vol.setDimensions(*sc.bounds); vol.setVolumeSize(*sc.dims); vol.constructVolume(); vol.integratePointCloud(sc.points, sc.normals)
    t0=time.perf_counter(); vol._commit(ctx); t1=time.perf_counter()
    ctx.lib.dmf_prepare_grid(ctx.h,1); t2=time.perf_counter()
    a=C.c_float(); b=C.c_float(); ctx.lib.dmf_volume_prepare_ms(ctx.h,C.byref(a),C.byref(b))
    print(name,"commit wall ms %.1f bytes wall ms %.1f build_ms %.2f bytes_ms %.2f"%(1e3*(t1-t0),1e3*(t2-t1),a.value,b.value), flush=True)
    ctx.close()
