# Dumps the reference's HJI cache (BicycleCAvoid.jld2: grid_knots, V_raw, ∇V_raw — src/HJI_computation.jl:47-51, deps/build.jl) into the
# flat "PGNHJI1" file that pigeon.jl_b200/hji_io.py and pgn_set_hji_cache read.  Run where Julia and JLD2.jl are installed:
#     julia export_hji_cache.jl BicycleCAvoid.jld2 BicycleCAvoid.pgnhji
using JLD2
src, dst = ARGS[1], ARGS[2]
@load src grid_knots V_raw ∇V_raw
dims = Int32[length(k) for k in grid_knots]
@assert size(V_raw) == Tuple(dims) && size(∇V_raw) == (7, dims...)
open(dst, "w") do io
    write(io, b"PGNHJI1\0")
    write(io, dims)                                  # little-endian on every platform Julia supports for this file
    for k in grid_knots
        write(io, convert(Vector{Float32}, k))
    end
    write(io, convert(Array{Float32}, V_raw))        # column-major: dimension 1 fastest
    write(io, convert(Array{Float32}, ∇V_raw))       # (7, n1, ..., n7): the 7 components fastest
end
println("wrote ", dst, " (", filesize(dst), " bytes, grid ", Tuple(dims), ")")
