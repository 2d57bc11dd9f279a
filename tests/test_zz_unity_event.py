"""Unity render-thread entry (include/gsplat_b200.h GsUnityFrameEvent): the callback a CommandBuffer's
IssuePluginEventAndData would run.  Named test_zz_* so it runs after the parity tests proper."""
import ctypes as C

import numpy as np
import pytest


def test_event_payload_layout_and_error_path():
    from unitygaussiansplatting_b200 import _native as N
    lib = N.native()
    assert lib.gs_unity_frame_event_size() == C.sizeof(N.GsUnityFrameEvent)
    assert N.GsUnityFrameEvent.params.offset == 16 and N.GsUnityFrameEvent.options.offset == 16 + C.sizeof(N.GsFrameParams)
    fn = lib.gs_unity_get_render_event_func()
    ev = N.GsUnityFrameEvent()
    for event_id in (N.GS_UNITY_EVENT_FRAME, N.GS_UNITY_EVENT_SYNC, 99):
        ev.status = -5
        fn(event_id, C.addressof(ev))           # null context: reported through `status`, nothing thrown, no CUDA call
        assert ev.status == -1
    fn(N.GS_UNITY_EVENT_FRAME, None)            # null payload is ignored


@pytest.mark.gpu
def test_event_frame_equals_direct_frame(g, ctx):
    from unitygaussiansplatting_b200 import _native as N
    from util import camera
    asset = g.synthetic_asset(g.SCENE_CLUSTERED, 20000, 0x5EED0051, "Medium")
    cam = camera(g, 200, 150)
    r = g.GaussianSplatRenderer(asset, ctx)
    direct = np.zeros((150, 200, 4), np.float16)
    r.SortAndRenderSplats(cam, rt=direct)
    r.ResetOrder()
    fp, keep = g.make_frame_params(cam, r.localToWorldMatrix, r.m_SplatScale, r.m_OpacityScale, r.m_SHOrder, r.m_SHOnly)
    via_event = np.zeros_like(direct)
    ev = N.GsUnityFrameEvent()
    ev.ctx, ev.asset = C.cast(ctx.handle, C.c_void_p).value, C.cast(r._asset, C.c_void_p).value
    C.memmove(C.byref(ev.params), C.byref(fp), C.sizeof(fp))
    ev.options.blend_mode = 0
    ev.do_sort, ev.status = 1, -5
    ev.has_rt, ev.has_camera_target = 1, 0
    ev.rt.data, ev.rt.width, ev.rt.height = via_event.ctypes.data, 200, 150
    ev.rt.row_pitch_bytes, ev.rt.format, ev.rt.memory = 0, 0, 0
    fn = N.native().gs_unity_get_render_event_func()
    fn(N.GS_UNITY_EVENT_FRAME, C.addressof(ev))
    assert ev.status == 0
    fn(N.GS_UNITY_EVENT_SYNC, C.addressof(ev))
    assert ev.status == 0
    assert direct.any() and np.array_equal(via_event.view(np.uint16), direct.view(np.uint16))
    r.Dispose()
