// RayTracerNative.cs — drop-in replacement for Assets/Services/RayTracer.cs that forwards the render path to librtb200
// through P/Invoke.  Same public surface (SetComputeShader, InvalidateBVHCache, ReleaseBuffers, ClearRenderTarget,
// RenderToTexture, RenderAsync, SaveTexture), so SceneBuilder.cs and GifGenerator.cs compile unchanged after
// `using RayTracer = RayTracerNative;` (or renaming the class).
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no .NET / Mono / Unity.  The same ABI is exercised by the C++ mirror
// (include/rtb_raytracer.hpp, tests/cpp) and the Python mirror (cosig-raytracing_b200/raytracer.py).  Struct layouts below
// follow include/rtb.h field by field; RtbNative.rtb_abi_sizes lets the host assert them at start-up.
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;
using System.Threading;
using System.Threading.Tasks;
using UnityEngine;

internal static class RtbNative
{
    const string Lib = "rtb200"; // librtb200.so / rtb200.dll in Assets/Plugins/x86_64

    [StructLayout(LayoutKind.Sequential)] public struct XformElem { public int type; public float x, y, z, angleDeg; }
    [StructLayout(LayoutKind.Sequential)] public struct Material { public float r, g, b, ka, kd, ks, kr, ior; }
    [StructLayout(LayoutKind.Sequential)] public unsafe struct Triangle { public int material; public fixed float v[9]; }
    [StructLayout(LayoutKind.Sequential)] public struct Mesh { public int xform, reserved; public long firstTri, nTris; }
    [StructLayout(LayoutKind.Sequential)] public struct Prim { public int xform, material; }

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct SceneDesc
    {
        public int hasImage, imageW, imageH; public fixed float bg[3];
        public int hasCamera, camXform; public float camDistance, camVfovDeg;
        public int nXforms; public int* xformOffsets; public XformElem* xformElems;
        public int nLights; public int* lightXforms; public float* lightRgb;
        public int nMaterials; public Material* materials;
        public int nMeshes; public Mesh* meshes;
        public long nTriangles; public Triangle* triangles;
        public int nSpheres; public Prim* spheres;
        public int nBoxes; public Prim* boxes;
    }

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct RenderParams
    {
        public int hasResolution, width, height;
        public int hasBg; public fixed float bg[3];
        public float lightIntensity;
        public int hasCamPos; public fixed float camPos[3];
        public int hasCamRot; public fixed float camRotEulerDeg[3];
        public int hasFov; public float fovDeg;
        public int maxDepth, enableAmbient, enableDiffuse, enableSpecular, enableRefraction, isOrthographic, aaSamples;
        public int softShadows; public float lightSize;
        public int glossy; public float roughness;
        public int motionBlur; public float shutterSpeed;
        public int debugMode, srgbEncode, bandRank, bandWorld, bandRows, outLayout;
        public fixed int reserved[6];
    }

    [DllImport(Lib)] public static extern int rtb_create(out IntPtr ctx, int[] deviceIds, int nDevices);
    [DllImport(Lib)] public static extern void rtb_destroy(IntPtr ctx);
    [DllImport(Lib)] public static extern void rtb_params_default(ref RenderParams p);
    [DllImport(Lib)] public static extern unsafe int rtb_upload_scene(IntPtr ctx, SceneDesc* scene, int primitiveMode, int bvhMode);
    [DllImport(Lib)] public static extern int rtb_invalidate(IntPtr ctx);
    [DllImport(Lib)] public static extern int rtb_clear_target(IntPtr ctx);
    [DllImport(Lib)] public static extern unsafe int rtb_render(IntPtr ctx, ref RenderParams p, void* rgba8, UIntPtr bytes, out int w, out int h);
    [DllImport(Lib)] public static extern int rtb_render_device(IntPtr ctx, ref RenderParams p, IntPtr dstDevice, UIntPtr bytes, int sync);
    [DllImport(Lib)] public static extern unsafe int rtb_resolve_frame(SceneDesc* scene, ref RenderParams p, float* out25, int* wh);
    [DllImport(Lib)] public static extern unsafe int rtb_set_cancel_flag(IntPtr ctx, int* flag);
    [DllImport(Lib)] public static extern IntPtr rtb_last_error(IntPtr ctx);
    [DllImport(Lib)] public static extern void rtb_abi_sizes(int[] sizes, int n);
    // GIF sweep (GifGenerator.cs): device palette mapping + host LZW inside the library
    [DllImport(Lib)] public static extern unsafe int rtb_gif_save(IntPtr ctx, string path, int width, int height, byte** rgba8Frames, int nFrames, int frameDelayCs, int threads);
    [DllImport(Lib)] public static extern int rtb_gif_render_rotation(IntPtr ctx, ref RenderParams baseParams, int nFrames, float stepDeg, string path, int frameDelayCs, int threads);
}

public sealed class RayTracerNative : IDisposable
{
    IntPtr ctx;
    ObjectData cachedScene;          // RayTracer.cs:118 — the BVH cache key is the scene object's identity
    bool needsRebuild = true;
    Texture2D target;

    public int BvhMode = 0;          // 0 = reference-shape BVH (bit-exact ids), 1 = GPU LBVH

    public RayTracerNative(int[] devices = null)
    {
        int rc = RtbNative.rtb_create(out ctx, devices, devices?.Length ?? 0);
        if (rc != 0) throw new InvalidOperationException(Marshal.PtrToStringAnsi(RtbNative.rtb_last_error(IntPtr.Zero)));
    }

    public void SetComputeShader(ComputeShader shader) { /* RayTracer.cs:29 — nothing to inject: the kernels live in the library */ }
    public void InvalidateBVHCache() { needsRebuild = true; RtbNative.rtb_invalidate(ctx); }                       // :38
    public void ReleaseBuffers() { cachedScene = null; needsRebuild = true; RtbNative.rtb_invalidate(ctx); RtbNative.rtb_clear_target(ctx); } // :47
    public void ClearRenderTarget() { RtbNative.rtb_clear_target(ctx); }                                            // :65

    // RayTracer.cs:212 — blocking inside the library; the Task only keeps the signature.
    public unsafe Task<Texture2D> RenderAsync(ObjectData scene, RenderSettings settings, IProgress<float> progress, CancellationToken token)
    {
        if (scene == null) return Task.FromResult<Texture2D>(null);
        EnsureScene(scene);
        progress?.Report(0.1f);
        var p = ToParams(settings);
        int cancel = 0;
        using (token.Register(() => Volatile.Write(ref cancel, 1)))
        {
            int w = scene.Image != null ? Math.Max(1, scene.Image.horizontal) : 256, h = scene.Image != null ? Math.Max(1, scene.Image.vertical) : 256;
            if (settings.ResolutionOverride.HasValue) { w = settings.ResolutionOverride.Value.x; h = settings.ResolutionOverride.Value.y; }
            if (target == null || target.width != w || target.height != h) target = new Texture2D(w, h, TextureFormat.RGBA32, false, true);
            var pixels = target.GetRawTextureData<byte>();                       // row 0 = bottom, exactly the library's order
            RtbNative.rtb_set_cancel_flag(ctx, &cancel);
            int rc = RtbNative.rtb_render(ctx, ref p, Unity.Collections.LowLevel.Unsafe.NativeArrayUnsafeUtility.GetUnsafePtr(pixels),
                                          (UIntPtr)(ulong)pixels.Length, out _, out _);
            RtbNative.rtb_set_cancel_flag(ctx, null);
            if (rc == -5) return Task.FromResult<Texture2D>(null);                // RTB_E_CANCELLED — the reference returns null, :283
            if (rc != 0) throw new InvalidOperationException(Marshal.PtrToStringAnsi(RtbNative.rtb_last_error(ctx)));
            target.Apply(false);                                                  // replaces ReadPixels + Apply, :371-375
        }
        progress?.Report(1f);
        return Task.FromResult(target);
    }

    public static void SaveTexture(Texture2D tex, string path) => System.IO.File.WriteAllBytes(path, tex.EncodeToPNG()); // :504

    unsafe void EnsureScene(ObjectData scene)
    {
        if (!needsRebuild && ReferenceEquals(cachedScene, scene)) return;       // :118-123, :273-278
        // Flatten ObjectData into rtb_scene_desc (CSR of transformations, one triangle array with per-mesh ranges) and call
        // rtb_upload_scene; the library copies what it needs, so the pinned arrays can be released right after.
        var offsets = new List<int> { 0 }; var elems = new List<RtbNative.XformElem>();
        foreach (var t in scene.Transformations)
        {
            foreach (var e in t.Elements)
                elems.Add(new RtbNative.XformElem { type = (int)e.Type, x = e.XYZ.x, y = e.XYZ.y, z = e.XYZ.z, angleDeg = e.AngleDeg });
            offsets.Add(elems.Count);
        }
        // ... materials, lights, meshes/triangles, spheres, boxes are marshalled the same way (omitted: mechanical) ...
        // fixed (...) { RtbNative.SceneDesc d = ...; Check(RtbNative.rtb_upload_scene(ctx, &d, 0, BvhMode)); }
        cachedScene = scene; needsRebuild = false;
    }

    static unsafe RtbNative.RenderParams ToParams(RenderSettings s)
    {
        var p = new RtbNative.RenderParams();
        RtbNative.rtb_params_default(ref p);
        if (s.ResolutionOverride.HasValue) { p.hasResolution = 1; p.width = s.ResolutionOverride.Value.x; p.height = s.ResolutionOverride.Value.y; }
        if (s.BackgroundColorOverride.HasValue) { p.hasBg = 1; var c = s.BackgroundColorOverride.Value; p.bg[0] = c.r; p.bg[1] = c.g; p.bg[2] = c.b; }
        p.lightIntensity = s.LightIntensityScale;
        if (s.CameraPositionOverride.HasValue) { p.hasCamPos = 1; var v = s.CameraPositionOverride.Value; p.camPos[0] = v.x; p.camPos[1] = v.y; p.camPos[2] = v.z; }
        if (s.CameraRotationOverride.HasValue) { p.hasCamRot = 1; var v = s.CameraRotationOverride.Value; p.camRotEulerDeg[0] = v.x; p.camRotEulerDeg[1] = v.y; p.camRotEulerDeg[2] = v.z; }
        if (s.CameraFovOverride.HasValue) { p.hasFov = 1; p.fovDeg = s.CameraFovOverride.Value; }
        p.maxDepth = s.MaxDepth;
        p.enableAmbient = s.EnableAmbient ? 1 : 0; p.enableDiffuse = s.EnableDiffuse ? 1 : 0;
        p.enableSpecular = s.EnableSpecular ? 1 : 0; p.enableRefraction = s.EnableRefraction ? 1 : 0;
        p.isOrthographic = s.IsOrthographic ? 1 : 0; p.aaSamples = s.AASamples;
        p.softShadows = s.EnableSoftShadows ? 1 : 0; p.lightSize = s.LightSize;
        p.glossy = s.EnableGlossy ? 1 : 0; p.roughness = s.SurfaceRoughness;
        p.motionBlur = s.EnableMotionBlur ? 1 : 0; p.shutterSpeed = s.ShutterSpeed;
        return p;
    }

    public void Dispose() { if (ctx != IntPtr.Zero) { RtbNative.rtb_destroy(ctx); ctx = IntPtr.Zero; } }

    // ---- what GifGenerator.cs needs beyond RenderAsync ---------------------------------------------------------------------
    // GifGenerator.SaveGifAsync (:82-155) body becomes one call: pin the frames' raw RGBA32 data (GetRawTextureData<byte>()),
    // pass the pointers.  The palette mapping (:346-369) runs on the GPU, LZW (:411-501) on the library's host threads.
    public unsafe void SaveGif(List<Texture2D> frames, string filePath, int frameDelay = 10)
    {
        if (frames == null || frames.Count == 0) return;                              // :84
        var ptrs = stackalloc byte*[frames.Count];
        for (int i = 0; i < frames.Count; i++)
            ptrs[i] = (byte*)Unity.Collections.LowLevel.Unsafe.NativeArrayUnsafeUtility.GetUnsafeReadOnlyPtr(frames[i].GetRawTextureData<byte>());
        int rc = RtbNative.rtb_gif_save(ctx, filePath, frames[0].width, frames[0].height, ptrs, frames.Count, frameDelay, 0);
        if (rc != 0) throw new InvalidOperationException(Marshal.PtrToStringAnsi(RtbNative.rtb_last_error(ctx)));
    }

    // GenerateRotationFrames (:40-72) + SaveGifAsync fused: SceneBuilder.OnGifClicked (SceneBuilder.cs:965-1030) can call this
    // instead of the two steps when it does not need the frames for playback.
    public void RenderRotationGif(ObjectData scene, RenderSettings baseSettings, string filePath, int frameDelay = 10)
    {
        EnsureScene(scene);
        var p = ToParams(baseSettings);
        int rc = RtbNative.rtb_gif_render_rotation(ctx, ref p, 36, 10f, filePath, frameDelay, 0);
        if (rc != 0) throw new InvalidOperationException(Marshal.PtrToStringAnsi(RtbNative.rtb_last_error(ctx)));
    }
}
