// RayTracerNative.cs — drop-in replacement for Assets/Services/RayTracer.cs that forwards the render path to librtb200
// through P/Invoke.  Same public surface (SetComputeShader :29, InvalidateBVHCache :38, ReleaseBuffers :47, ClearRenderTarget :65,
// RenderToTexture :82, RenderAsync :212, SaveTexture :504), so SceneBuilder.cs and GifGenerator.cs compile unchanged after
// `using RayTracer = RayTracerNative;` (or renaming the class).  Additions the library offers beyond that surface: RenderBegin /
// RenderEnd (pipelined frames), GetStats, SaveGif / RenderRotationGif, and an optional zero-copy bridge for RenderToTexture.
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no .NET / Mono / Unity.  What IS exercised here is the call sequence this
// file makes, statement for statement, by tests/c/binding_sequence.c (plain C against the same header), and the same ABI by the
// C++ mirror (include/rtb_raytracer.hpp) and the Python mirror (cosig-raytracing_b200/raytracer.py).  Struct layouts below follow
// include/rtb.h field by field; the static constructor asserts their sizes against rtb_abi_sizes at start-up.
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;
using System.Threading;
using System.Threading.Tasks;
using Unity.Collections;
using Unity.Collections.LowLevel.Unsafe;
using UnityEngine;

public static class RtbNative
{
    const string Lib = "rtb200"; // librtb200.so / rtb200.dll in Assets/Plugins/x86_64

    public const int OK = 0, E_ARG = -1, E_CUDA = -2, E_NOSCENE = -3, E_SIZE = -4, E_CANCELLED = -5, E_IO = -6, E_PARSE = -7;
    public const int PRIM_TESSELLATED = 0, PRIM_ANALYTIC = 1, BVH_REFERENCE = 0, BVH_LBVH = 1;
    public const int EXT_OPAQUE_FD = 1, EXT_OPAQUE_WIN32 = 2, EXT_D3D12_HEAP = 4, EXT_D3D12_RESOURCE = 5;

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

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct Stats
    {
        public long raysPrimary, raysContinuation, raysShadow, pathsHitPrimary, nTriangles, nNodes;
        public int width, height, spp, chunks, kernelLaunches, nDevices;
        public float msUpload, msBuild, msRenderDevice, msTraverse, msShade, msResolve;
        public long h2dBytes, d2hBytes;
        public fixed long reserved[4];
        public long raysTraversed, packetNodeFetches, packetTriFetches, bytesPerSlot;
    }

    [DllImport(Lib)] public static extern int rtb_api_version();
    [DllImport(Lib)] public static extern void rtb_abi_sizes(int[] sizes, int n);
    [DllImport(Lib)] public static extern int rtb_create(out IntPtr ctx, int[] deviceIds, int nDevices);
    [DllImport(Lib)] public static extern void rtb_destroy(IntPtr ctx);
    [DllImport(Lib)] public static extern void rtb_params_default(ref RenderParams p);
    [DllImport(Lib)] public static extern unsafe int rtb_upload_scene(IntPtr ctx, SceneDesc* scene, int primitiveMode, int bvhMode);
    [DllImport(Lib)] public static extern int rtb_invalidate(IntPtr ctx);
    [DllImport(Lib)] public static extern int rtb_clear_target(IntPtr ctx);
    [DllImport(Lib)] public static extern unsafe int rtb_render(IntPtr ctx, ref RenderParams p, void* rgba8, UIntPtr bytes, out int w, out int h);
    [DllImport(Lib)] public static extern unsafe int rtb_render_begin(IntPtr ctx, ref RenderParams p, void* rgba8, UIntPtr bytes, out int ticket);
    [DllImport(Lib)] public static extern int rtb_render_end(IntPtr ctx, int ticket);
    [DllImport(Lib)] public static extern int rtb_render_device(IntPtr ctx, ref RenderParams p, IntPtr dstDevice, UIntPtr bytes, int sync);
    [DllImport(Lib)] public static extern int rtb_external_import(IntPtr ctx, int handleType, IntPtr handle, UIntPtr bytes, int dedicated, out IntPtr devPtr);
    [DllImport(Lib)] public static extern int rtb_external_release(IntPtr ctx, IntPtr devPtr);
    [DllImport(Lib)] public static extern unsafe int rtb_resolve_frame(SceneDesc* scene, ref RenderParams p, float* out25, int* wh);
    [DllImport(Lib)] public static extern int rtb_set_cancel_flag(IntPtr ctx, IntPtr flag);
    [DllImport(Lib)] public static extern int rtb_get_stats(IntPtr ctx, out Stats stats);
    [DllImport(Lib)] public static extern int rtb_synchronize(IntPtr ctx);
    [DllImport(Lib)] public static extern IntPtr rtb_last_error(IntPtr ctx);
    [DllImport(Lib)] public static extern IntPtr rtb_alloc_pinned(UIntPtr bytes);
    [DllImport(Lib)] public static extern void rtb_free_pinned(IntPtr p);
    // GIF sweep (GifGenerator.cs): device palette mapping + host LZW inside the library
    [DllImport(Lib)] public static extern unsafe int rtb_gif_save(IntPtr ctx, string path, int width, int height, byte** rgba8Frames, int nFrames, int frameDelayCs, int threads);
    [DllImport(Lib)] public static extern int rtb_gif_render_rotation(IntPtr ctx, ref RenderParams baseParams, int nFrames, float stepDeg, string path, int frameDelayCs, int threads);

    static unsafe RtbNative()
    {
        var sizes = new int[9];
        rtb_abi_sizes(sizes, 9);
        int[] mine = { sizeof(XformElem), sizeof(Material), sizeof(Triangle), sizeof(Mesh), sizeof(Prim), sizeof(SceneDesc), sizeof(RenderParams), sizeof(Stats) };
        for (int i = 0; i < mine.Length; i++)
            if (mine[i] != sizes[i]) throw new InvalidOperationException($"rtb200 ABI mismatch: struct #{i} is {mine[i]} bytes here, {sizes[i]} in the library");
    }
}

/// <summary>Optional zero-copy bridge for RenderToTexture.  Unity does not hand out shareable memory handles from C#; a native
/// rendering plugin (IUnityGraphicsVulkan / IUnityGraphicsD3D12) does.  An implementation allocates the GraphicsBuffer the display
/// blit reads from in exportable memory and returns its handle (Vulkan: vkGetMemoryFdKHR -> RTB_EXT_OPAQUE_FD; D3D12:
/// CreateSharedHandle on the committed resource -> RTB_EXT_D3D12_RESOURCE, dedicated = true).  Without a bridge RenderToTexture
/// uses the pipelined host path (one PCIe round trip per frame) and still returns the same reused RenderTexture.</summary>
public interface IRtbTextureBridge
{
    /// <summary>Creates (or returns the cached) width*height*4-byte buffer shared with CUDA.</summary>
    bool TryGetSharedBuffer(int width, int height, out GraphicsBuffer buffer, out int handleType, out IntPtr handle, out bool dedicated);
    /// <summary>Copies the buffer's RGBA8 pixels (row 0 = bottom) into the texture on the graphics queue (a one-dispatch compute blit).</summary>
    void BlitToTexture(GraphicsBuffer buffer, RenderTexture target);
}

public sealed class RayTracerNative : IDisposable
{
    IntPtr ctx;
    ObjectData cachedScene;          // RayTracer.cs:118 — the BVH cache key is the scene object's identity
    bool needsRebuild = true;
    Texture2D target;                // RenderAsync's result, reused
    RenderTexture targetTexture;     // RenderToTexture's result, reused (RayTracer.cs:126-132: "do not destroy it")
    Texture2D staging;               // host path of RenderToTexture
    IntPtr cancelFlag;               // unmanaged int the library polls between wavefront depths
    IntPtr sharedDevPtr; int sharedW, sharedH; GraphicsBuffer sharedBuffer;

    public int BvhMode = RtbNative.BVH_REFERENCE;        // 0 = reference-shape BVH (bit-exact ids), 1 = GPU LBVH
    public int PrimitiveMode = RtbNative.PRIM_TESSELLATED;
    public IRtbTextureBridge TextureBridge;              // null: host path

    public RayTracerNative(int[] devices = null)
    {
        int rc = RtbNative.rtb_create(out ctx, devices, devices?.Length ?? 0);
        if (rc != RtbNative.OK) throw new InvalidOperationException(Marshal.PtrToStringAnsi(RtbNative.rtb_last_error(IntPtr.Zero)));
        cancelFlag = Marshal.AllocHGlobal(4);
        Marshal.WriteInt32(cancelFlag, 0);
    }

    void Check(int rc) { if (rc != RtbNative.OK) throw new InvalidOperationException($"rtb200 error {rc}: {Marshal.PtrToStringAnsi(RtbNative.rtb_last_error(ctx))}"); }

    public void SetComputeShader(ComputeShader shader) { /* RayTracer.cs:29 — nothing to inject: the kernels live in the library */ }
    public void InvalidateBVHCache() { needsRebuild = true; Check(RtbNative.rtb_invalidate(ctx)); }                                   // :38
    public void ReleaseBuffers() { cachedScene = null; needsRebuild = true; Check(RtbNative.rtb_invalidate(ctx)); ClearRenderTarget(); } // :47
    public void ClearRenderTarget()                                                                                                    // :65
    {
        ReleaseShared();
        Check(RtbNative.rtb_clear_target(ctx));
        if (targetTexture != null) { targetTexture.Release(); targetTexture = null; }
    }

    static void Resolve(ObjectData scene, RenderSettings settings, out int w, out int h)   // RayTracer.cs:93-94 / :221-222
    {
        w = settings.ResolutionOverride.HasValue ? settings.ResolutionOverride.Value.x : Mathf.Max(1, scene.Image != null ? scene.Image.horizontal : 256);
        h = settings.ResolutionOverride.HasValue ? settings.ResolutionOverride.Value.y : Mathf.Max(1, scene.Image != null ? scene.Image.vertical : 256);
    }

    // RayTracer.cs:212 — blocking inside the library; the Task only keeps the signature.
    public unsafe Task<Texture2D> RenderAsync(ObjectData scene, RenderSettings settings, IProgress<float> progress, CancellationToken token)
    {
        if (scene == null) return Task.FromResult<Texture2D>(null);
        EnsureScene(scene);
        progress?.Report(0.1f);
        if (token.IsCancellationRequested) return Task.FromResult<Texture2D>(null);                 // :283
        var p = ToParams(settings);
        Resolve(scene, settings, out int w, out int h);
        if (target == null || target.width != w || target.height != h) target = new Texture2D(w, h, TextureFormat.RGBA32, false, true);
        NativeArray<byte> pixels = target.GetRawTextureData<byte>();                               // row 0 = bottom, exactly the library's order
        Marshal.WriteInt32(cancelFlag, 0);
        IntPtr flag = cancelFlag;
        int rc;
        using (token.Register(() => Marshal.WriteInt32(flag, 1)))
        {
            Check(RtbNative.rtb_set_cancel_flag(ctx, cancelFlag));
            rc = RtbNative.rtb_render(ctx, ref p, NativeArrayUnsafeUtility.GetUnsafePtr(pixels), (UIntPtr)(ulong)pixels.Length, out _, out _);
            RtbNative.rtb_set_cancel_flag(ctx, IntPtr.Zero);
        }
        if (rc == RtbNative.E_CANCELLED) return Task.FromResult<Texture2D>(null);                   // the reference returns null, :283
        Check(rc);
        target.Apply(false);                                                                        // replaces ReadPixels + Apply, :371-375
        progress?.Report(1f);
        return Task.FromResult(target);
    }

    // RayTracer.cs:82-202 — the realtime path: returns the reused GPU RenderTexture, never a CPU copy of it.
    public unsafe RenderTexture RenderToTexture(ObjectData scene, RenderSettings settings)
    {
        if (scene == null) return null;
        EnsureScene(scene);
        var p = ToParams(settings);
        Resolve(scene, settings, out int w, out int h);
        if (targetTexture == null || targetTexture.width != w || targetTexture.height != h)        // :126-132
        {
            if (targetTexture != null) targetTexture.Release();
            targetTexture = new RenderTexture(w, h, 0, RenderTextureFormat.ARGB32) { enableRandomWrite = true };
            targetTexture.Create();
        }
        ulong bytes = (ulong)w * (ulong)h * 4UL;
        if (TextureBridge != null && EnsureShared(w, h, bytes))
        {
            // zero copy: k_resolve stores the pixels into memory Unity's graphics device owns; sync = 1 orders them before the blit
            Check(RtbNative.rtb_render_device(ctx, ref p, sharedDevPtr, (UIntPtr)bytes, 1));
            TextureBridge.BlitToTexture(sharedBuffer, targetTexture);
            return targetTexture;
        }
        // host path: pipelined begin / end into the staging texture's own memory, then one GPU blit
        if (staging == null || staging.width != w || staging.height != h) staging = new Texture2D(w, h, TextureFormat.RGBA32, false, true);
        NativeArray<byte> pixels = staging.GetRawTextureData<byte>();
        Check(RtbNative.rtb_render_begin(ctx, ref p, NativeArrayUnsafeUtility.GetUnsafePtr(pixels), (UIntPtr)bytes, out int ticket));
        Check(RtbNative.rtb_render_end(ctx, ticket));
        staging.Apply(false);
        Graphics.Blit(staging, targetTexture);
        return targetTexture;
    }

    bool EnsureShared(int w, int h, ulong bytes)
    {
        if (sharedDevPtr != IntPtr.Zero && sharedW == w && sharedH == h) return true;
        ReleaseShared();
        if (!TextureBridge.TryGetSharedBuffer(w, h, out sharedBuffer, out int type, out IntPtr handle, out bool dedicated)) return false;
        int rc = RtbNative.rtb_external_import(ctx, type, handle, (UIntPtr)bytes, dedicated ? 1 : 0, out sharedDevPtr);
        if (rc != RtbNative.OK) { sharedDevPtr = IntPtr.Zero; Debug.LogWarning("rtb200: external memory import failed, using the host path: " + Marshal.PtrToStringAnsi(RtbNative.rtb_last_error(ctx))); return false; }
        sharedW = w; sharedH = h;
        return true;
    }

    void ReleaseShared()
    {
        if (sharedDevPtr != IntPtr.Zero) { RtbNative.rtb_external_release(ctx, sharedDevPtr); sharedDevPtr = IntPtr.Zero; }
        sharedBuffer = null;
    }

    // Pipelined RenderAsync for callers that render a stream of frames (the GIF sweep, an offline turntable): up to 16 tickets in
    // flight; `pixels` (e.g. a Texture2D's raw data or a persistent NativeArray) must stay valid until RenderEnd(ticket).
    public unsafe int RenderBegin(ObjectData scene, RenderSettings settings, NativeArray<byte> pixels)
    {
        EnsureScene(scene);
        var p = ToParams(settings);
        Check(RtbNative.rtb_render_begin(ctx, ref p, NativeArrayUnsafeUtility.GetUnsafePtr(pixels), (UIntPtr)(ulong)pixels.Length, out int ticket));
        return ticket;
    }
    public void RenderEnd(int ticket) { Check(RtbNative.rtb_render_end(ctx, ticket)); }

    public RtbNative.Stats GetStats() { Check(RtbNative.rtb_get_stats(ctx, out RtbNative.Stats s)); return s; }

    public static void SaveTexture(Texture2D tex, string path)                                                          // :504-509
    {
        byte[] png = tex.EncodeToPNG();
        System.IO.Directory.CreateDirectory(System.IO.Path.GetDirectoryName(path));
        System.IO.File.WriteAllBytes(path, png);
    }

    // RebuildBVH (RayTracer.cs:386-404) + SetupMaterialBuffer (:455-499): flatten ObjectData into rtb_scene_desc — CSR of the
    // transformations, one triangle array with a range per mesh — and hand it over.  The library copies what it needs during the
    // call, so everything is pinned only for its duration.  The cache flags change only after the upload succeeded.
    unsafe void EnsureScene(ObjectData scene)
    {
        if (!needsRebuild && ReferenceEquals(cachedScene, scene)) return;       // :118-123, :273-278
        var offsets = new List<int> { 0 };
        var elems = new List<RtbNative.XformElem>();
        foreach (var t in scene.Transformations)
        {
            foreach (var e in t.Elements)
                elems.Add(new RtbNative.XformElem { type = (int)e.Type, x = e.XYZ.x, y = e.XYZ.y, z = e.XYZ.z, angleDeg = e.AngleDeg });
            offsets.Add(elems.Count);
        }
        var lightXforms = new int[scene.Lights.Count];
        var lightRgb = new float[scene.Lights.Count * 3];
        for (int i = 0; i < scene.Lights.Count; i++)
        {
            lightXforms[i] = scene.Lights[i].transformationIndex;
            lightRgb[3 * i] = scene.Lights[i].rgb.r; lightRgb[3 * i + 1] = scene.Lights[i].rgb.g; lightRgb[3 * i + 2] = scene.Lights[i].rgb.b;
        }
        var materials = new RtbNative.Material[scene.Materials.Count];
        for (int i = 0; i < materials.Length; i++)
        {
            var m = scene.Materials[i];
            materials[i] = new RtbNative.Material { r = m.color.r, g = m.color.g, b = m.color.b, ka = m.ambient, kd = m.diffuse, ks = m.specular, kr = m.refraction, ior = m.ior };
        }
        long nTris = 0;
        foreach (var mesh in scene.TriangleMeshes) nTris += mesh.Triangles.Count;
        var meshes = new RtbNative.Mesh[scene.TriangleMeshes.Count];
        var tris = new RtbNative.Triangle[nTris];
        long at = 0;
        for (int i = 0; i < meshes.Length; i++)
        {
            var mesh = scene.TriangleMeshes[i];
            meshes[i] = new RtbNative.Mesh { xform = mesh.transformationIndex, reserved = 0, firstTri = at, nTris = mesh.Triangles.Count };
            foreach (var t in mesh.Triangles)
            {
                RtbNative.Triangle o = default;
                o.material = t.materialIndex;
                o.v[0] = t.v0.x; o.v[1] = t.v0.y; o.v[2] = t.v0.z;
                o.v[3] = t.v1.x; o.v[4] = t.v1.y; o.v[5] = t.v1.z;
                o.v[6] = t.v2.x; o.v[7] = t.v2.y; o.v[8] = t.v2.z;
                tris[at++] = o;
            }
        }
        var spheres = new RtbNative.Prim[scene.Spheres.Count];
        for (int i = 0; i < spheres.Length; i++) spheres[i] = new RtbNative.Prim { xform = scene.Spheres[i].transformationIndex, material = scene.Spheres[i].materialIndex };
        var boxes = new RtbNative.Prim[scene.Boxes.Count];
        for (int i = 0; i < boxes.Length; i++) boxes[i] = new RtbNative.Prim { xform = scene.Boxes[i].transformationIndex, material = scene.Boxes[i].materialIndex };
        int[] offsetArr = offsets.ToArray();
        RtbNative.XformElem[] elemArr = elems.ToArray();

        fixed (int* pOffsets = offsetArr, pLightX = lightXforms)
        fixed (float* pLightRgb = lightRgb)
        fixed (RtbNative.XformElem* pElems = elemArr)
        fixed (RtbNative.Material* pMats = materials)
        fixed (RtbNative.Mesh* pMeshes = meshes)
        fixed (RtbNative.Triangle* pTris = tris)
        fixed (RtbNative.Prim* pSpheres = spheres, pBoxes = boxes)
        {
            RtbNative.SceneDesc d = default;
            if (scene.Image != null)
            {
                d.hasImage = 1; d.imageW = scene.Image.horizontal; d.imageH = scene.Image.vertical;
                d.bg[0] = scene.Image.background.r; d.bg[1] = scene.Image.background.g; d.bg[2] = scene.Image.background.b;
            }
            if (scene.Camera != null)
            {
                d.hasCamera = 1; d.camXform = scene.Camera.transformationIndex;
                d.camDistance = scene.Camera.distance; d.camVfovDeg = scene.Camera.verticalFovDeg;
            }
            d.nXforms = scene.Transformations.Count; d.xformOffsets = pOffsets; d.xformElems = pElems;
            d.nLights = lightXforms.Length; d.lightXforms = pLightX; d.lightRgb = pLightRgb;
            d.nMaterials = materials.Length; d.materials = pMats;
            d.nMeshes = meshes.Length; d.meshes = pMeshes;
            d.nTriangles = nTris; d.triangles = pTris;
            d.nSpheres = spheres.Length; d.spheres = pSpheres;
            d.nBoxes = boxes.Length; d.boxes = pBoxes;
            Check(RtbNative.rtb_upload_scene(ctx, &d, PrimitiveMode, BvhMode));
        }
        cachedScene = scene;
        needsRebuild = false;
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

    public void Dispose()
    {
        if (ctx == IntPtr.Zero) return;
        ReleaseShared();
        RtbNative.rtb_destroy(ctx);
        ctx = IntPtr.Zero;
        if (cancelFlag != IntPtr.Zero) { Marshal.FreeHGlobal(cancelFlag); cancelFlag = IntPtr.Zero; }
        if (targetTexture != null) { targetTexture.Release(); targetTexture = null; }
    }

    // ---- what GifGenerator.cs needs beyond RenderAsync ---------------------------------------------------------------------
    // GifGenerator.SaveGifAsync (:82-155) body becomes one call: pin the frames' raw RGBA32 data (GetRawTextureData<byte>()),
    // pass the pointers.  The palette mapping (:346-369) runs on the GPU, LZW (:411-501) on the library's host threads.
    public unsafe void SaveGif(List<Texture2D> frames, string filePath, int frameDelay = 10)
    {
        if (frames == null || frames.Count == 0) return;                              // :84
        var ptrs = stackalloc byte*[frames.Count];
        for (int i = 0; i < frames.Count; i++)
            ptrs[i] = (byte*)NativeArrayUnsafeUtility.GetUnsafeReadOnlyPtr(frames[i].GetRawTextureData<byte>());
        Check(RtbNative.rtb_gif_save(ctx, filePath, frames[0].width, frames[0].height, ptrs, frames.Count, frameDelay, 0));
    }

    // GenerateRotationFrames (:40-72) + SaveGifAsync fused: SceneBuilder.OnGifClicked (SceneBuilder.cs:965-1030) can call this
    // instead of the two steps when it does not need the frames for playback.
    public void RenderRotationGif(ObjectData scene, RenderSettings baseSettings, string filePath, int frameDelay = 10)
    {
        EnsureScene(scene);
        var p = ToParams(baseSettings);
        Check(RtbNative.rtb_gif_render_rotation(ctx, ref p, 36, 10f, filePath, frameDelay, 0));
    }
}
