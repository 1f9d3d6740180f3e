# Fused (outer-boundary) route of INTEGRATION.md section 2, second bullet: no file of Clarabel.jl
# is touched.  `include` this after B200Ext.jl (inside the same package extension).  It adds
# more-specific methods for the two internal functions that already receive the concrete LDL
# engine as an argument (kktsolver_directldl.jl:211 and :389) plus a fused `solve!`, so that with
# `Settings(direct_solve_method = :b200)` the cone -> K update, the static regularisation, the
# factorisation, the triangular solves and the iterative refinement all run on the device and only
# the cone state, the right-hand sides and the solutions cross PCIe.
#
# NOT EXECUTED IN THIS REPOSITORY (no Julia in the build image).  The executable twin is
# clarabel.jl_b200/kktsolver_b200.py::B200KKTSolver, which issues exactly the same C calls and is
# what tests/test_gpu_*.py and bench.py run.

import Clarabel: _kktsolver_update_inner!, _iterative_refinement, DirectLDLKKTSolver, CompositeCone,
                 ZeroCone, NonnegativeCone, SecondOrderCone, PSDTriangleCone, ExponentialCone,
                 PowerCone, GenPowerCone, SOCExpansionMap, GenPowExpansionMap, numel, pack_triu

# per-engine state of the fused route (kept in a side table so that B200DirectLDLSolver stays the
# plain inner-boundary object of B200Ext.jl)
mutable struct FusedState
    n::Int; m::Int
    w::Vector{Float64}; eta::Vector{Float64}; d::Vector{Float64}
    u::Vector{Float64}; v::Vector{Float64}; R::Vector{Float64}
    ns_index::Vector{Int64}; ns_values::Vector{Float64}      # nonsymmetric cones: update_values! payload
    last_status::Int32
end
const FUSED = IdDict{B200DirectLDLSolver,FusedState}()

_cone_code(::ZeroCone) = Int32(0);        _cone_code(::NonnegativeCone) = Int32(1)
_cone_code(::SecondOrderCone) = Int32(2); _cone_code(::PSDTriangleCone) = Int32(3)
_cone_code(::ExponentialCone) = Int32(4); _cone_code(::PowerCone) = Int32(5)
_cone_code(::GenPowerCone) = Int32(6)
_cone_dim(K::PSDTriangleCone) = Int64(K.n)
_cone_dim(K) = Int64(numel(K))

function _fused_setup(ks::DirectLDLKKTSolver{Float64}, ldl::B200DirectLDLSolver{Float64}, cones)
    map = ks.map
    ctype = Int32[_cone_code(K) for K in cones]
    cdim  = Int64[_cone_dim(K) for K in cones]
    soc_u = Int64[]; soc_v = Int64[]; soc_D = Int64[]
    gp = Int64[]                                   # genpow q, r, p, D positions, cone by cone
    for sm in map.sparse_maps
        if sm isa SOCExpansionMap
            append!(soc_u, sm.u); append!(soc_v, sm.v); append!(soc_D, sm.D)
        end
    end
    # positions owned by the nonsymmetric cones (order = order of the values in _fused_nonsym!)
    ns_index = Int64[]
    for (K, rng) in zip(cones, cones.rng_blocks)
        (K isa ExponentialCone || K isa PowerCone || K isa GenPowerCone) && append!(ns_index, map.Hsblocks[rng])
    end
    for sm in map.sparse_maps
        sm isa GenPowExpansionMap && append!(gp, sm.q)
    end
    for sm in map.sparse_maps
        sm isa GenPowExpansionMap && append!(gp, sm.r)
    end
    for sm in map.sparse_maps
        sm isa GenPowExpansionMap && append!(gp, sm.p)
    end
    for sm in map.sparse_maps
        sm isa GenPowExpansionMap && append!(gp, sm.D)
    end
    append!(ns_index, gp)
    rc = ccall((:cb200_set_maps, LIB), Int32,
        (Ptr{Cvoid}, Int64, Int64, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Int64, Ptr{Int64}, Int64,
         Ptr{Int64}, Int64, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
        ldl.handle, ks.n, ks.m, ks.p, map.P, length(map.P), map.A, length(map.A),
        map.Hsblocks, length(map.Hsblocks), map.diag_full, length(ctype), ctype, cdim,
        soc_u, soc_v, soc_D)
    _chk(rc)
    nsoc  = count(K -> K isa SecondOrderCone, cones)
    nsrow = sum(K -> K isa SecondOrderCone ? numel(K) : 0, cones; init = 0)
    nR    = sum(K -> K isa PSDTriangleCone ? K.n^2 : 0, cones; init = 0)
    FUSED[ldl] = FusedState(ks.n, ks.m, zeros(ks.m), zeros(nsoc), zeros(nsoc), zeros(nsrow),
                            zeros(nsrow), zeros(nR), ns_index, zeros(length(ns_index)), Int32(0))
end

# gather the scaling state `update_scaling!` just produced (SURVEY.md H5: upload the state, derive
# Hs on the device)
function _fused_state!(st::FusedState, cones)
    isoc = 0; irow = 0; iR = 0
    for (K, rng) in zip(cones, cones.rng_cones)
        if K isa NonnegativeCone
            st.w[rng] .= K.w
        elseif K isa SecondOrderCone
            isoc += 1; dim = numel(K)
            st.w[rng] .= K.w; st.eta[isoc] = K.η
            if !isnothing(K.sparse_data)
                st.d[isoc] = K.sparse_data.d
                st.u[irow+1:irow+dim] .= K.sparse_data.u
                st.v[irow+1:irow+dim] .= K.sparse_data.v
            end
            irow += dim
        elseif K isa PSDTriangleCone
            nn = K.n^2
            st.R[iR+1:iR+nn] .= vec(K.data.R); iR += nn
        end
    end
end

# values of the entries the nonsymmetric cones own, exactly as get_Hs! / _csc_update_sparsecone
# (directldl_datamaps.jl:146-166) would write them
function _fused_nonsym!(st::FusedState, cones)
    k = 0
    blk6 = zeros(6)
    for K in cones
        if K isa ExponentialCone || K isa PowerCone
            pack_triu(blk6, K.Hs); st.ns_values[k+1:k+6] .= .-blk6; k += 6
        elseif K isa GenPowerCone
            d = K.data; d1n = length(d.d1); dn = numel(K)
            st.ns_values[k+1:k+d1n] .= .-d.μ .* d.d1
            st.ns_values[k+d1n+1:k+dn] .= -d.μ * d.d2
            k += dn
        end
    end
    for field in (:q, :r, :p)
        for K in cones
            if K isa GenPowerCone
                vsrc = getfield(K.data, field); len = length(vsrc)
                st.ns_values[k+1:k+len] .= .-sqrt(K.data.μ) .* vsrc; k += len
            end
        end
    end
    for K in cones
        if K isa GenPowerCone
            st.ns_values[k+1:k+3] .= (-1.0, -1.0, 1.0); k += 3
        end
    end
end

# kktsolver_update! -> here (kktsolver_directldl.jl:196-208 re-dispatches on the engine type)
function _kktsolver_update_inner!(ks::DirectLDLKKTSolver{Float64}, ldl::B200DirectLDLSolver{Float64},
                                  cones::CompositeCone{Float64})
    haskey(FUSED, ldl) || _fused_setup(ks, ldl, cones)
    st = FUSED[ldl]
    if !isempty(st.ns_index)
        _fused_nonsym!(st, cones)
        update_values!(ldl, st.ns_index, st.ns_values)
    end
    _fused_state!(st, cones)
    rc = ccall((:cb200_update_cones, LIB), Int32,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        ldl.handle, st.w, st.eta, st.d, st.u, st.v, st.R)
    return _chk(rc)          # false = numerical failure, as _kktsolver_regularize_and_refactor! reports it
end

# kktsolver_solve! calls solve!(ldlsolver, KKT, x, b) and then _iterative_refinement (:346-371):
# the fused solve does both on the device; b = [rhsx; rhsz; 0_p], x = [lhsx; lhsz; .]
function solve!(ldl::B200DirectLDLSolver{Float64}, K::SparseMatrixCSC{Float64}, x::Vector{Float64}, b::Vector{Float64})
    st = get(FUSED, ldl, nothing)
    if isnothing(st)         # plain inner-boundary use (update_values!/refactor! driven by Julia)
        return _chk(ccall((:cb200_solve, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), ldl.handle, x, b))
    end
    rounds = Ref{Int32}(0)
    GC.@preserve x b begin
        st.last_status = ccall((:cb200_solve_ir, LIB), Int32,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
            ldl.handle, pointer(b), pointer(b, st.n + 1), pointer(x), pointer(x, st.n + 1), rounds)
    end
    st.last_status < 0 && error(unsafe_string(ccall((:cb200_last_error, LIB), Cstring, ())))
    return nothing
end

# the refinement already happened inside cb200_solve_ir; report its verdict
function _iterative_refinement(ks::DirectLDLKKTSolver{Float64}, ldl::B200DirectLDLSolver{Float64})
    st = get(FUSED, ldl, nothing)
    isnothing(st) && return invoke(_iterative_refinement,
                                   Tuple{DirectLDLKKTSolver{Float64},Clarabel.AbstractDirectLDLSolver{Float64}}, ks, ldl)
    return st.last_status == 0
end
