# Dumps the reference's HJI cache (BicycleCAvoid.jld2: grid_knots, V_raw, ∇V_raw — src/HJI_computation.jl:47-51, deps/build.jl) into the
# flat "PGNHJI1" file that pigeon.jl_b200/hji_io.py and pgn_set_hji_cache read.  (pigeon.jl_b200/jld2.py also reads the .jld2 container
# directly; this dump is the fallback that does not depend on the container format.)  Run where Julia and JLD2.jl are installed:
#     julia export_hji_cache.jl BicycleCAvoid.jld2 BicycleCAvoid.pgnhji
using JLD2
src, dst = ARGS[1], ARGS[2]
@load src grid_knots V_raw ∇V_raw
dims = Int32[length(k) for k in grid_knots]
# save() stores Array(reinterpret(Float32, cache.∇V.coefs)) (src/HJI_computation.jl:59-64): its size is (7*n1, n2, ..., n7) on Julia 1.0.x and
# (7, n1, ..., n7) on later versions — the memory order (7 components fastest, then dimension 1) is the same, only the length is checked
@assert size(V_raw) == Tuple(dims) && length(∇V_raw) == 7 * prod(dims)
open(dst, "w") do io
    write(io, b"PGNHJI1\0")
    write(io, dims)                                  # little-endian on every platform Julia supports for this file
    for k in grid_knots
        write(io, convert(Vector{Float32}, k))
    end
    write(io, convert(Array{Float32}, V_raw))        # column-major: dimension 1 fastest
    write(io, convert(Vector{Float32}, vec(∇V_raw)))  # the 7 components fastest, then dimension 1
end
println("wrote ", dst, " (", filesize(dst), " bytes, grid ", Tuple(dims), ")")
