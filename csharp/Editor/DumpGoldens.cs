// DumpGoldens.cs — Unity EDITOR script that pins this repository's oracle against the real reference.
//
// The reference (mpoboas/cosig-raytracing) ships no golden images and cannot run in the build image (Unity C# + HLSL), so
// oracle/oracle.cpp is "parity unpinned" (DESIGN.md §2).  A maintainer with the reference project open in Unity 6000.0.61f1 closes
// that gap with this file: copy it to Assets/Editor/, run  Tools > rtb200 > Dump golden frames,  and copy the produced folder to
// this repository's tests/golden/unity/.  tests/test_unity_goldens.py then compares the oracle (and, on a GPU, librtb200) with what
// the STOCK RayTracer + BVHRayTracing.compute rendered: final RGBA8 frames, and the kernel's own depth debug view (_DebugMode 1,
// BVHRayTracing.compute:484-497), which exposes the primary hit distance t of every pixel at 8-bit resolution.
//
// What is rendered: the three shipped scenes (Assets/Resources/Scenes/{test_scene_1,test_scene_2,eval_scene}.txt) at BASELINE
// config C1 — 320x240, depth 3, all lighting toggles on, intensity 1, no camera overrides (the M_scene.inverse branch,
// RayTracer.cs:266), 1 and 4 samples per pixel — through RayTracer.RenderToTexture, i.e. exactly the reference's code path.
// Nothing of the reference is modified.  The depth view re-dispatches the same kernel with only _DebugMode changed: every buffer
// and uniform is still bound on the ComputeShader from the RenderToTexture call, and its Result target is the RenderTexture that
// call returned.
//
// Output (row 0 = bottom, Texture2D order; 4 bytes per pixel):
//   <scene>_aa<N>.rgba, <scene>_depth.rgba, manifest.json (resolution, settings, Unity version, graphics API, colour space).
#if UNITY_EDITOR
using System.Collections.Generic;
using System.IO;
using System.Text;
using UnityEditor;
using UnityEngine;

public static class DumpGoldens
{
    const int Width = 320, Height = 240, Depth = 3;
    static readonly string[] Scenes = { "test_scene_1", "test_scene_2", "eval_scene" };

    [MenuItem("Tools/rtb200/Dump golden frames")]
    public static void Dump()
    {
        string outDir = EditorUtility.SaveFolderPanel("Folder for the golden frames", "", "rtb200_goldens");
        if (string.IsNullOrEmpty(outDir)) return;
        var shader = AssetDatabase.LoadAssetAtPath<ComputeShader>("Assets/Shaders/BVHRayTracing.compute");
        if (shader == null) { Debug.LogError("Assets/Shaders/BVHRayTracing.compute not found: open the reference project"); return; }
        var tracer = new RayTracer();
        tracer.SetComputeShader(shader);
        var entries = new List<string>();
        foreach (string name in Scenes)
        {
            ObjectData scene = new SceneService().LoadScene(Path.Combine(Application.dataPath, "Resources/Scenes/" + name + ".txt"));
            foreach (int aa in new[] { 1, 4 })
            {
                RenderSettings s = Settings(aa);
                RenderTexture rt = tracer.RenderToTexture(scene, s);
                string file = $"{name}_aa{aa}.rgba";
                File.WriteAllBytes(Path.Combine(outDir, file), ReadBack(rt));
                entries.Add(Entry(name, file, aa, 0));
                if (aa == 1)
                {
                    // same bindings, same target; only the debug switch differs (the host always passes 0, RayTracer.cs:143)
                    int kernel = shader.FindKernel("CSMain");
                    shader.SetInt("_DebugMode", 1);
                    shader.Dispatch(kernel, Mathf.CeilToInt(Width / 8.0f), Mathf.CeilToInt(Height / 8.0f), 1);
                    shader.SetInt("_DebugMode", 0);
                    string dfile = $"{name}_depth.rgba";
                    File.WriteAllBytes(Path.Combine(outDir, dfile), ReadBack(rt));
                    entries.Add(Entry(name, dfile, 1, 1));
                }
            }
            tracer.InvalidateBVHCache();
        }
        tracer.ReleaseBuffers();
        var sb = new StringBuilder();
        sb.Append("{\n \"unity_version\": \"").Append(Application.unityVersion).Append("\",\n");
        sb.Append(" \"graphics_api\": \"").Append(SystemInfo.graphicsDeviceType).Append("\",\n");
        sb.Append(" \"graphics_device\": \"").Append(SystemInfo.graphicsDeviceName.Replace("\"", "'")).Append("\",\n");
        sb.Append(" \"color_space\": \"").Append(QualitySettings.activeColorSpace).Append("\",\n");
        sb.Append(" \"width\": ").Append(Width).Append(", \"height\": ").Append(Height).Append(", \"max_depth\": ").Append(Depth).Append(",\n");
        sb.Append(" \"frames\": [\n  ").Append(string.Join(",\n  ", entries)).Append("\n ]\n}\n");
        File.WriteAllText(Path.Combine(outDir, "manifest.json"), sb.ToString());
        Debug.Log($"rtb200: wrote {entries.Count} frames to {outDir}");
    }

    static RenderSettings Settings(int aa)
    {
        // GetRenderSettingsFromUI's values for a static render (SceneBuilder.cs:404-487) minus the camera overrides, so the
        // headless branch cameraToObject = M_scene.inverse (RayTracer.cs:266) is the one exercised, as in this repository's C1
        return new RenderSettings
        {
            ResolutionOverride = new Vector2Int(Width, Height),
            LightIntensityScale = 1.0f,
            MaxDepth = Depth,
            EnableAmbient = true, EnableDiffuse = true, EnableSpecular = true, EnableRefraction = true,
            IsOrthographic = false,
            AASamples = aa,
            EnableSoftShadows = false, LightSize = 0f, EnableGlossy = false, SurfaceRoughness = 0f, EnableMotionBlur = false, ShutterSpeed = 0f,
        };
    }

    static byte[] ReadBack(RenderTexture rt)
    {
        // the reference's own readback (RayTracer.cs:371-375): ReadPixels into an RGBA32 Texture2D
        var prev = RenderTexture.active;
        RenderTexture.active = rt;
        var tex = new Texture2D(rt.width, rt.height, TextureFormat.RGBA32, false, true);
        tex.ReadPixels(new Rect(0, 0, rt.width, rt.height), 0, 0);
        tex.Apply();
        RenderTexture.active = prev;
        byte[] raw = tex.GetRawTextureData();
        Object.DestroyImmediate(tex);
        return raw;
    }

    static string Entry(string scene, string file, int aa, int debug) =>
        $"{{\"scene\": \"{scene}\", \"file\": \"{file}\", \"aa_samples\": {aa}, \"debug_mode\": {debug}}}";
}
#endif
