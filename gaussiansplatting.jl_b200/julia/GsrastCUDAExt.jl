# GsrastCUDAExt.jl — reference-side binding of libgsrast.so (include/gsrast.h).
#
# What a GaussianSplatting.jl maintainer adds to ext/GaussianSplattingCUDAExt to route the rasterizer hot path
# (`rasterize` / `∇rasterize` / `_update_stats!`, src/rasterization/rasterizer.jl:255-550, src/strategy.jl:107-136)
# through the sm_100a library instead of the KernelAbstractions kernels.  The `GaussianRasterizer` struct, the
# functor's activation pre-pass (rasterizer.jl:200-253), the `rrule` (rasterizer.jl:552-573) and the image / GL
# helpers stay untouched: only the bodies of the three functions below change.
#
# NOT EXECUTED IN THIS REPOSITORY: the build image has no Julia toolchain.  It mirrors, call for call, the ctypes
# binding in gsrast/_lib.py + gsrast/rasterizer.py, which the test-suite does exercise.
module GsrastCUDAExt

using CUDA
using StaticArrays
import GaussianSplatting as GSP
import GaussianSplatting: GaussianRasterizer, Camera, n_color_features

const libgsrast = get(ENV, "GSRAST_LIB", "libgsrast.so")

# ---- C structs (include/gsrast.h) --------------------------------------------------------------------------
struct GsrConfig
    width::Int32; height::Int32; channels::Int32
    near_plane::Float32; far_plane::Float32
    radius_clip::Int32; blur_eps::Float32; math_mode::Int32
end

struct GsrCamera
    R::NTuple{9, Float32}          # column-major, i.e. Tuple(SMatrix{3,3,Float32})
    t::NTuple{3, Float32}
    focal::NTuple{2, Float32}
    principal::NTuple{2, Float32}
    cam_center::NTuple{3, Float32}
    R_dev::CuPtr{Float32}          # C_NULL unless R_w2c / t_w2c are passed positionally (pose optimisation)
    t_dev::CuPtr{Float32}
end

struct GsrStateViews
    n::Int64; n_rendered::Int64
    radii::CuPtr{Int32}; grad_means2d::CuPtr{Float32}; means2d::CuPtr{Float32}; depths::CuPtr{Float32}
    conics::CuPtr{Float32}; rgbs::CuPtr{Float32}; clamped::CuPtr{UInt8}; tiles_touched::CuPtr{Int32}
    points_offset::CuPtr{Int32}; normals::CuPtr{Float32}; keys_unsorted::CuPtr{UInt64}
    values_unsorted::CuPtr{UInt32}; keys_sorted::CuPtr{UInt64}; values_sorted::CuPtr{UInt32}
    ranges::CuPtr{UInt32}; n_contrib::CuPtr{UInt32}; accum_alpha::CuPtr{Float32}
end

const GSR_MATH_REFERENCE = Int32(0)
const GSR_MATH_FAST = Int32(1)
const GSR_MATH_STRICT = Int32(2)    # default: flat 1e-5 / 1e-4 parity with the reference

function check(rc::Cint, h::Ptr{Cvoid} = C_NULL)
    rc == 0 && return
    msg = unsafe_string(ccall((:gsr_last_error, libgsrast), Cstring, (Ptr{Cvoid},), h))
    error("libgsrast error $rc: $msg")          # same surface as the reference's `error(...)` / `@assert`
end

# One native handle per GaussianRasterizer (two coexist with a sky dome, src/sky_dome.jl:143-145).
const HANDLES = IdDict{GaussianRasterizer, Ptr{Cvoid}}()

function handle(rast::GaussianRasterizer)
    get!(HANDLES, rast) do
        width, height = size(rast.image, 2), size(rast.image, 3)
        cfg = Ref(GsrConfig(width, height, n_color_features(rast.mode), rast.near_plane, rast.far_plane,
                            Int32(3), 0.3f0, GSR_MATH_STRICT))   # radius_clip, blur_ϵ: rasterizer.jl:294-295
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:gsr_create, libgsrast), Cint, (Ref{GsrConfig}, Ref{Ptr{Cvoid}}), cfg, out))
        finalizer(r -> ccall((:gsr_destroy, libgsrast), Cint, (Ptr{Cvoid},), pop!(HANDLES, r, C_NULL)), rast)
        out[]
    end
end

function c_camera(camera::Camera, R_w2c, t_w2c)
    K = camera.intrinsics
    R = SMatrix{3, 3, Float32}(camera.w2c[1:3, 1:3]); t = SVector{3, Float32}(camera.w2c[1:3, 4])
    GsrCamera(Tuple(R), Tuple(t), Tuple(K.focal), Tuple(K.principal), Tuple(camera.camera_center),
              R_w2c ≡ nothing ? CU_NULL : pointer(R_w2c), t_w2c ≡ nothing ? CU_NULL : pointer(t_w2c))
end

maybe_ptr(x) = x ≡ nothing ? CU_NULL : pointer(x)
stream_ptr() = CUDA.stream().handle

# ---- rasterize — replaces the kernel launches of rasterizer.jl:283-407 -------------------------------------
function GSP.rasterize(
    means_3d::CuMatrix{Float32}, shs::CuArray{Float32, 3}, opacities::CuMatrix{Float32},
    scales::CuMatrix{Float32}, rotations::CuMatrix{Float32}, R_w2c = nothing, t_w2c = nothing;
    rast::GaussianRasterizer, camera::Camera, sh_degree::Int, background::SVector{3, Float32},
    covisibilities = nothing, uncertainties = nothing,
)
    h = handle(rast)
    n, K = size(means_3d, 2), size(shs, 2)
    cam = Ref(c_camera(camera, R_w2c, t_w2c))
    bg = Ref(Tuple(background))
    n_rendered = Ref{Int64}(0)
    GC.@preserve means_3d shs opacities scales rotations R_w2c t_w2c covisibilities uncertainties begin
        check(ccall((:gsr_forward, libgsrast), Cint,
            (Ptr{Cvoid}, Ref{GsrCamera}, Int64, Int32, Int32, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
             CuPtr{Float32}, CuPtr{Float32}, Ref{NTuple{3, Float32}}, CuPtr{Float32}, CuPtr{Bool}, CuPtr{Float32},
             Ref{Int64}, Ptr{Cvoid}),
            h, cam, n, sh_degree, K, pointer(means_3d), pointer(shs), pointer(opacities), pointer(scales),
            pointer(rotations), bg, pointer(rast.image), maybe_ptr(covisibilities), maybe_ptr(uncertainties),
            n_rendered, stream_ptr()), h)
    end
    refresh_state_views!(rast, h, n)
    return rast.image
end

# ---- ∇rasterize — replaces rasterizer.jl:437-549 ------------------------------------------------------------
function GSP.∇rasterize(
    vpixels::CuArray{Float32, 3}, means_3d::CuMatrix{Float32}, shs::CuArray{Float32, 3},
    scales::CuMatrix{Float32}, rotations::CuMatrix{Float32}, opacities::CuMatrix{Float32},
    radii, R_w2c = nothing, t_w2c = nothing;
    rast::GaussianRasterizer, camera::Camera, sh_degree::Int, background::SVector{3, Float32},
)
    h = handle(rast)
    n, K = size(means_3d, 2), size(shs, 2)
    # outputs need no zero-fill: the library writes every row (zeros for culled Gaussians)
    vmeans = CuArray{Float32}(undef, 3, n); vshs = CuArray{Float32}(undef, size(shs))
    vopacities = CuArray{Float32}(undef, 1, n); vscales = CuArray{Float32}(undef, 3, n)
    vrot = CuArray{Float32}(undef, 4, n)
    vR = R_w2c ≡ nothing ? nothing : CUDA.zeros(Float32, 3, 3)     # rasterizer.jl:500-501
    vt = R_w2c ≡ nothing ? nothing : CUDA.zeros(Float32, 3)
    cam = Ref(c_camera(camera, R_w2c, t_w2c)); bg = Ref(Tuple(background))
    GC.@preserve vpixels means_3d shs scales rotations opacities R_w2c t_w2c begin
        check(ccall((:gsr_backward, libgsrast), Cint,
            (Ptr{Cvoid}, Ref{GsrCamera}, Int64, Int32, Int32, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
             CuPtr{Float32}, CuPtr{Float32}, Ref{NTuple{3, Float32}}, CuPtr{Float32}, CuPtr{Float32},
             CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int32,
             Ptr{Cvoid}),
            h, cam, n, sh_degree, K, pointer(means_3d), pointer(shs), pointer(opacities), pointer(scales),
            pointer(rotations), bg, pointer(vpixels), pointer(vmeans), pointer(vshs), pointer(vopacities),
            pointer(vscales), pointer(vrot), maybe_ptr(vR), maybe_ptr(vt), Int32(0), stream_ptr()), h)
    end
    return vmeans, vshs, vopacities, vscales, vrot, vR, vt
end

# rast.gstate.radii / .∇means_2d are read by the densification strategy (strategy.jl:85-86) and the debug report
# (training.jl:562): re-point them at the library-owned buffers after every forward.
function refresh_state_views!(rast::GaussianRasterizer, h, n)
    v = Ref{GsrStateViews}()
    check(ccall((:gsr_get_state, libgsrast), Cint, (Ptr{Cvoid}, Ref{GsrStateViews}), h, v), h)
    radii = unsafe_wrap(CuArray, v[].radii, n)
    ∇means_2d = unsafe_wrap(CuArray, reinterpret(CuPtr{SVector{2, Float32}}, v[].grad_means2d), n)
    rast.gstate = GSP.GeometryState(rast.gstate; radii, ∇means_2d)   # a keyword re-constructor, 5 lines in states.jl
    return
end

# ---- update_stats! — replaces the `_update_stats!` launch of strategy.jl:107-116 ----------------------------
function GSP.update_stats!(strategy::GSP.DefaultStrategy, rast::GaussianRasterizer)
    h = handle(rast)
    check(ccall((:gsr_update_stats, libgsrast), Cint,
        (Ptr{Cvoid}, Int64, CuPtr{Int32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
        h, length(strategy.max_radii), pointer(strategy.max_radii), pointer(strategy.accum_∇means_2d),
        pointer(strategy.denom), stream_ptr()), h)
end

# ---- fused SSIM — replaces the two kernel launches of fused_ssim.jl:354-389; the rrule (:397-407) is unchanged ----
function GSP._fused_ssim(img::CuArray{Float32, 4}; ref::CuArray{Float32, 4}, C1::Float32 = 0.01f0^2,
                         C2::Float32 = 0.03f0^2, train::Bool)
    W, H, CH, B = size(img)
    ssim_map = CuArray{Float32}(undef, W, H, CH, B)
    d = ntuple(_ -> train ? CuArray{Float32}(undef, W, H, CH, B) : CuArray{Float32}(undef, 0, 0, 0, 0), 3)
    check(ccall((:gsr_ssim_forward, libgsrast), Cint,
        (Int32, Int32, Int32, Int32, CuPtr{Float32}, CuPtr{Float32}, Float32, Float32, Int32, CuPtr{Float32},
         CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
        W, H, CH, B, pointer(img), pointer(ref), C1, C2, train, pointer(ssim_map),
        train ? pointer(d[1]) : CU_NULL, train ? pointer(d[2]) : CU_NULL, train ? pointer(d[3]) : CU_NULL, stream_ptr()))
    return ssim_map, d...
end

function GSP.fused_ssim_bwd(img::T, ref::T, dL_dmap::T, dm_dmu1::T, dm_dsigma1_sq::T, dm_dsigma12::T;
                            C1::Float32 = 0.01f0^2, C2::Float32 = 0.03f0^2) where T <: CuArray{Float32, 4}
    W, H, CH, B = size(img)
    dL_dimg = CuArray{Float32}(undef, W, H, CH, B)      # every element is written: no zero-fill
    check(ccall((:gsr_ssim_backward, libgsrast), Cint,
        (Int32, Int32, Int32, Int32, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
        W, H, CH, B, pointer(img), pointer(ref), pointer(dL_dmap), pointer(dm_dmu1), pointer(dm_dsigma1_sq),
        pointer(dm_dsigma12), pointer(dL_dimg), stream_ptr()))
    return dL_dimg
end

# The activation-fused functor (gsr_forward_raw / gsr_backward_raw) binds the same way as rasterize / ∇rasterize above:
# one `rrule` on `(rast::GaussianRasterizer)(means_3d, opacities, scales, rotations, sh_color, sh_remainder; ...)` whose
# forward calls :gsr_forward_raw and whose pullback calls :gsr_backward_raw (argument order in include/gsrast.h).

GSP.release_scene_buffers!(rast::GaussianRasterizer) =
    check(ccall((:gsr_release_scene_buffers, libgsrast), Cint, (Ptr{Cvoid},), handle(rast)))

function GSP.memory_usage(rast::GaussianRasterizer)
    b = Ref{Csize_t}(0)
    check(ccall((:gsr_memory_usage, libgsrast), Cint, (Ptr{Cvoid}, Ref{Csize_t}), handle(rast), b))
    Int(b[]) + sizeof(rast.image) + sizeof(rast.shs) + sizeof(rast.scales_act) + sizeof(rast.opacities_act)
end

end # module
