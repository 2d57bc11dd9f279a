// GaussianSplatNativeFrame.cs -- Unity-side shim for libgsplat_b200 (SURVEY 8f row N2).
//
// NOT compiled or run in this repository (no Unity / dotnet toolchain in the build image): it is the file a maintainer
// drops next to package/Runtime/GaussianSplatRenderer.cs, together with GaussianSplatNative.cs from INTEGRATION.md.
// What it replaces: the per-renderer body of GaussianSplatRenderSystem.SortAndRenderSplats
// (package/Runtime/GaussianSplatRenderer.cs:118-166: SortPoints + CalcViewData + DrawProcedural) by ONE native event on
// Unity's render thread.  The native side of the event (gs_unity_get_render_event_func, GsUnityFrameEvent in
// include/gsplat_b200.h) is built and tested here (tests/test_zz_unity_event.py).
using System;
using System.Runtime.InteropServices;
using Unity.Collections;
using Unity.Collections.LowLevel.Unsafe;
using UnityEngine;
using UnityEngine.Rendering;

namespace GaussianSplatting.Runtime
{
    /// One of these per GaussianSplatRenderer; owns the pinned event payloads handed to the render thread.
    internal sealed unsafe class GaussianSplatNativeFrame : IDisposable
    {
        [StructLayout(LayoutKind.Sequential)]
        struct GsUnityFrameEvent                                    // include/gsplat_b200.h (size checked against gs_unity_frame_event_size)
        {
            public IntPtr ctx, asset;
            public GaussianSplatNative.GsFrameParams fp;
            public GaussianSplatNative.GsRenderOptions options;
            public int do_sort, status;
            public uint has_rt, has_camera_target;
            public GaussianSplatNative.GsImage rt, camera_target;
        }

        [DllImport("gsplat_b200")] static extern IntPtr gs_unity_get_render_event_func();
        [DllImport("gsplat_b200")] static extern uint gs_unity_frame_event_size();

        const int kEventFrame = 1;                                  // GS_UNITY_EVENT_FRAME
        const int kInFlight = 3;                                    // command buffers may execute up to 2 frames late
        NativeArray<GsUnityFrameEvent> m_Events;                    // persistent => stable address, no GC pinning needed
        // The reference binds _SplatCutouts / _SplatDeletedBits / _SplatSelectedBits on every CalcViewData (SetAssetDataOnCS,
        // R/GaussianSplatRenderer.cs:485-509).  The native frame takes them as HOST arrays borrowed until the event has run,
        // so every in-flight slot owns its copies.  The bit arrays live in GraphicsBuffers in the reference (edited by compute
        // kernels); a host that edits must keep a CPU mirror of them (hostDeletedBits / hostSelectedBits below) -- without
        // one, pass default and the frame renders as if nothing were deleted or selected.
        NativeArray<GaussianCutout.ShaderData>[] m_Cutouts = new NativeArray<GaussianCutout.ShaderData>[kInFlight];
        NativeArray<uint>[] m_Deleted = new NativeArray<uint>[kInFlight], m_Selected = new NativeArray<uint>[kInFlight];
        int m_Next;
        static readonly IntPtr s_EventFunc = gs_unity_get_render_event_func();

        public GaussianSplatNativeFrame()
        {
            if (gs_unity_frame_event_size() != (uint)UnsafeUtility.SizeOf<GsUnityFrameEvent>())
                throw new InvalidOperationException("GsUnityFrameEvent layout mismatch between C# and libgsplat_b200");
            m_Events = new NativeArray<GsUnityFrameEvent>(kInFlight, Allocator.Persistent);
        }

        public void Dispose()
        {
            if (m_Events.IsCreated) m_Events.Dispose();
            for (int i = 0; i < kInFlight; ++i)
            {
                if (m_Cutouts[i].IsCreated) m_Cutouts[i].Dispose();
                if (m_Deleted[i].IsCreated) m_Deleted[i].Dispose();
                if (m_Selected[i].IsCreated) m_Selected[i].Dispose();
            }
        }

        static void CopyInto<T>(ref NativeArray<T> dst, NativeArray<T> src) where T : unmanaged
        {
            if (!dst.IsCreated || dst.Length != src.Length)
            {
                if (dst.IsCreated) dst.Dispose();
                dst = new NativeArray<T>(src.Length, Allocator.Persistent, NativeArrayOptions.UninitializedMemory);
            }
            dst.CopyFrom(src);
        }

        /// Records the native frame into `cmb`.  `rtDevicePtr` is the CUDA mapping of the linear RGBA16F buffer that backs
        /// _GaussianSplatRT (external-memory interop, INTEGRATION.md section 3); the composite draw that follows in the
        /// reference (:206-210) stays as it is and reads that buffer.
        public void Record(CommandBuffer cmb, GaussianSplatRenderer gs, Camera cam, IntPtr ctx, IntPtr asset, IntPtr rtDevicePtr,
                           int rtWidth, int rtHeight, NativeArray<uint> hostDeletedBits = default, NativeArray<uint> hostSelectedBits = default)
        {
            // results of the event recorded kInFlight frames ago: log and skip, like the reference (:655)
            var ev = (GsUnityFrameEvent*)m_Events.GetUnsafePtr() + m_Next;
            if (ev->ctx != IntPtr.Zero && ev->status != 0 && ev->status != -5)
                GaussianSplatNative.Check(ev->ctx, ev->status);

            var tr = gs.transform;
            ev->ctx = ctx;
            ev->asset = asset;
            ev->fp = new GaussianSplatNative.GsFrameParams
            {
                mat_object_to_world = tr.localToWorldMatrix,
                mat_world_to_object = tr.worldToLocalMatrix,
                mat_view = cam.worldToCameraMatrix,
                mat_proj_gpu = GL.GetGPUProjectionMatrix(cam.projectionMatrix, true),   // UNITY_MATRIX_P is an engine global
                screen_w = rtWidth, screen_h = rtHeight,
                cam_pos_world = cam.transform.position,
                splat_scale = gs.m_SplatScale, opacity_scale = gs.m_OpacityScale,
                sh_order = (uint)gs.m_SHOrder, sh_only = gs.m_SHOnly ? 1u : 0u,
            };
            // cutouts: the ShaderData records UpdateCutoutsBuffer builds (R/GaussianSplatRenderer.cs:511-531), matrix = cutout
            // worldToLocal * renderer localToWorld (R/GaussianCutout.cs:26-40)
            int slot = m_Next;
            ev->fp.cutouts = null; ev->fp.cutout_count = 0; ev->fp.deleted_bits = null; ev->fp.selected_bits = null;
            if (gs.m_Cutouts != null && gs.m_Cutouts.Length > 0)
            {
                var data = new NativeArray<GaussianCutout.ShaderData>(gs.m_Cutouts.Length, Allocator.Temp);
                for (int i = 0; i < gs.m_Cutouts.Length; ++i) data[i] = GaussianCutout.GetShaderData(gs.m_Cutouts[i], tr.localToWorldMatrix);
                CopyInto(ref m_Cutouts[slot], data);
                data.Dispose();
                ev->fp.cutouts = (GaussianCutout.ShaderData*)m_Cutouts[slot].GetUnsafePtr();
                ev->fp.cutout_count = (uint)m_Cutouts[slot].Length;
            }
            int words = (gs.splatCount + 31) / 32;
            if (hostDeletedBits.IsCreated && hostDeletedBits.Length >= words)
            {
                CopyInto(ref m_Deleted[slot], hostDeletedBits);
                ev->fp.deleted_bits = (uint*)m_Deleted[slot].GetUnsafePtr();
            }
            if (hostSelectedBits.IsCreated && hostSelectedBits.Length >= words)
            {
                CopyInto(ref m_Selected[slot], hostSelectedBits);
                ev->fp.selected_bits = (uint*)m_Selected[slot].GetUnsafePtr();
            }
            ev->options = default;                                                    // fp16-ROP blend, whole image
            ev->do_sort = gs.m_FrameCounter % gs.m_SortNthFrame == 0 ? 1 : 0;         // :120-121
            ev->status = -5;                                                          // GS_ERR_NOT_READY until it has run
            ev->has_rt = 1; ev->has_camera_target = 0;
            ev->rt = new GaussianSplatNative.GsImage
            {
                data = (void*)rtDevicePtr, width = (uint)rtWidth, height = (uint)rtHeight,
                row_pitch_bytes = 0, format = 0 /* GS_PIX_RGBA16F */, memory = 1 /* GS_MEM_DEVICE */
            };
            ++gs.m_FrameCounter;
            cmb.IssuePluginEventAndData(s_EventFunc, kEventFrame, (IntPtr)ev);
            m_Next = (m_Next + 1) % kInFlight;
        }
    }
}
