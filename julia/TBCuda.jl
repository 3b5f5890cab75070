# TBCuda.jl -- reference-side binding of libtbcuda.so (UNTESTED in this repository's image: no Julia
# toolchain is available there; written against include/tbcuda.h).
#
# Drop-in for the hot path of TensorBranching.jl:
#     TensorBranching.solve_slice(branch, element_type, usecuda)      src/dynamic_ob.jl:30-34
#     TensorBranching.contract_slices(branches, element_type, usecuda) src/dynamic_ob.jl:36-48
# With `using TBCuda`, calls with usecuda=true are routed to the B200 engine; usecuda=false keeps the
# reference's own CPU route (TropicalGEMM).
module TBCuda

using TensorBranching
using TensorBranching: SlicedBranch, CompressedEinsum
using OMEinsum.OMEinsumContractionOrders: ContractionTree
using Graphs: nv

const LIB = get(ENV, "LIBTBCUDA", "libtbcuda.so")

# mirrors `struct tb_options` / `struct tb_network` of include/tbcuda.h
struct TbOptions
    device::Int32; n_devices::Int32; arena_bytes::Int64
    max_wave::Int32; host_threads::Int32; plan_flags::UInt32; streams_per_device::Int32
    devices::Ptr{Int32}; slice_budget::Int32; timing::Int32
end
struct TbNetwork
    n_labels::Int32; n_leaves::Int32
    leaf_off::Ptr{Int32}; leaf_labels::Ptr{Int32}
    n_open::Int32; open_labels::Ptr{Int32}
    node_left::Ptr{Int32}; node_right::Ptr{Int32}
    weights::Ptr{Cvoid}; weight_dtype::Int32; value_type::Int32
    flags::UInt32; n_fixed::Int32
    fixed_labels::Ptr{Int32}; fixed_values::Ptr{UInt8}   # index slicing (tb_contract_sliced fills these itself)
end

const CTX = Ref{Ptr{Cvoid}}(C_NULL)

# One engine per Julia process.  `devices` = the GPUs to use: with more than one, libtbcuda shards every
# contract_slices call over them itself (LPT by tropical ops) and combines with ONE ncclAllReduce(max) -- no
# Distributed.jl needed (tb_init_multi).  `slice_budget` > 0 lets it cut a call with fewer branches than 2 x GPUs into
# 2^k index slices.  Set TBCUDA_DEVICES="0,1,2,3,4,5,6,7" to pick the devices without touching code.
function ctx(devices::AbstractVector{<:Integer} = parse.(Int, split(get(ENV, "TBCUDA_DEVICES", "0"), ",")); slice_budget::Integer = 6)
    if CTX[] == C_NULL
        devs = Int32.(devices)
        opts = Ref(TbOptions(devs[1], 0, 0, 0, 0, 0, 0, C_NULL, slice_budget, 0))
        rc = GC.@preserve devs ccall((:tb_init_multi, LIB), Cint, (Ptr{Int32}, Int32, Ref{TbOptions}, Ref{Ptr{Cvoid}}),
                                     devs, length(devs), opts, CTX)
        rc == 0 || error("tb_init_multi failed ($rc): " * unsafe_string(ccall((:tb_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        atexit(() -> ccall((:tb_shutdown, LIB), Cint, (Ptr{Cvoid},), CTX[]))
    end
    return CTX[]
end

# ContractionTree (leaves = 1-based tensor ids) -> 0-based post-order child arrays
function flatten_tree(ct, n_leaves::Int)
    left = Int32[]; right = Int32[]
    function walk(t)
        t isa Integer && return Int32(t - 1)
        l = walk(t.left); r = walk(t.right)
        push!(left, l); push!(right, r)
        return Int32(n_leaves + length(left) - 1)
    end
    walk(ct)
    return left, right
end

weight_code(::Type{Int32}) = Int32(1); weight_code(::Type{Int64}) = Int32(2)
weight_code(::Type{Float32}) = Int32(3); weight_code(::Type{Float64}) = Int32(4)

struct FlatBranch   # keeps the arrays alive while the C call runs
    leaf_off::Vector{Int32}; leaf_labels::Vector{Int32}; left::Vector{Int32}; right::Vector{Int32}
    open::Vector{Int32}; weights::Any; net::TbNetwork
end

# tb_value_type: 0 = AUTO (exact integers for unit / integer weights, Tropical{Float32} for real weights),
# 4 = Tropical{Float64} (element_type Float64 with real weights; slower: no tiled GEMM kernel for 8-byte values)
value_type_for(::Type{T}, w) where {T} = (T === Float64 && !integer_valued(w)) ? Int32(4) : Int32(0)

function FlatBranch(branch::SlicedBranch, element_type::Type = Float32)
    code = branch.code::CompressedEinsum
    ixs = code.ixs
    leaf_off = Int32[0]; leaf_labels = Int32[]
    for ix in ixs
        append!(leaf_labels, Int32.(ix .- 1)); push!(leaf_off, Int32(length(leaf_labels)))
    end
    left, right = length(ixs) == 1 ? (Int32[], Int32[]) : flatten_tree(code.ct, length(ixs))
    open = Int32.(code.iy .- 1)
    w = branch.p.weights
    unit = w isa TensorBranching.UnitWeight
    wv = unit ? nothing : collect(w)
    net = TbNetwork(nv(branch.p.g), length(ixs), pointer(leaf_off), pointer(leaf_labels), length(open),
                    isempty(open) ? C_NULL : pointer(open), isempty(left) ? C_NULL : pointer(left),
                    isempty(right) ? C_NULL : pointer(right), unit ? C_NULL : Ptr{Cvoid}(pointer(wv)),
                    unit ? Int32(0) : weight_code(eltype(wv)), value_type_for(element_type, w), UInt32(0), Int32(0), C_NULL, C_NULL)
    return FlatBranch(leaf_off, leaf_labels, left, right, open, wv, net)
end

"contract_slices on the B200 engine: one C call for the whole branch list."
function contract_slices_cuda(branches::Vector{SlicedBranch}, element_type::Type)
    n = length(branches)
    flats = Vector{Union{FlatBranch, Nothing}}(undef, n)
    nets = Vector{TbNetwork}(undef, n)
    empty_net = TbNetwork(0, 0, C_NULL, C_NULL, 0, C_NULL, C_NULL, C_NULL, C_NULL, 0, 0, 0, 0, C_NULL, C_NULL)
    for (i, b) in enumerate(branches)
        if nv(b.p.g) == 0 || isnothing(b.code)
            flats[i] = nothing; nets[i] = empty_net
        else
            flats[i] = FlatBranch(b, element_type); nets[i] = flats[i].net
        end
    end
    vals = Vector{Float64}(undef, n); status = Vector{Int32}(undef, n); mx = Ref{Float64}(0)
    GC.@preserve flats begin
        rc = ccall((:tb_contract_networks, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{TbNetwork}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Int32}, Ref{Float64}),
                   ctx(), nets, C_NULL, n, vals, status, mx)
        rc == 0 || error("tb_contract_networks failed ($rc): " *
                         unsafe_string(ccall((:tb_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx())))
    end
    # same arithmetic as src/dynamic_ob.jl:39-44: empty graph => element_type(r), else t + element_type(r)
    return element_type[(nv(b.p.g) == 0 || isnothing(b.code)) ? element_type(b.r) : element_type(vals[i]) + element_type(b.r)
                        for (i, b) in enumerate(branches)]
end

"""
One heavy branch as 2^k index slices (tb_suggest_slices + tb_contract_sliced): the same value as solve_slice, from
2^k independent contractions of the same tree.  `range` = the assignments this process contracts (a multi-GPU job
gives every rank a disjoint range and combines the maxima with one all-reduce(max)).
"""
function solve_slice_index_sliced(branch::SlicedBranch, element_type::Type, k::Integer; range = nothing)
    flat = FlatBranch(branch)
    labels = Vector{Int32}(undef, k); sc = Ref{Float64}(0); tc = Ref{Float64}(0)
    GC.@preserve flat begin
        nk = ccall((:tb_suggest_slices, LIB), Cint,
                   (Ptr{Cvoid}, Ref{TbNetwork}, Int32, Int32, Ptr{Int32}, Ref{Float64}, Ref{Float64}),
                   C_NULL, Ref(flat.net), -1, k, labels, sc, tc)
        nk >= 0 || error("tb_suggest_slices failed ($nk)")
        a0, cnt = isnothing(range) ? (0, 1 << nk) : (Base.first(range), length(range))
        mx = Ref{Float64}(0)
        rc = ccall((:tb_contract_sliced, LIB), Cint,
                   (Ptr{Cvoid}, Ref{TbNetwork}, Ptr{Int32}, Int32, Int64, Int64, Float64, Ptr{Float64}, Ptr{Int32}, Ref{Float64}),
                   ctx(), Ref(flat.net), labels, nk, a0, cnt, 0.0, C_NULL, C_NULL, mx)
        rc == 0 || error("tb_contract_sliced failed ($rc): " *
                         unsafe_string(ccall((:tb_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx())))
    end
    return element_type(mx[])
end

# ---- streaming hand-off (tb_stream_*): push the finished slices of every round of slice_bfs / slice_dfs
#      (src/slice.jl:79-86) while the host keeps slicing; finish returns what contract_slices returns for all of them
mutable struct BranchStream
    handle::Ptr{Cvoid}
    branches::Vector{SlicedBranch}
end

function stream_begin(capacity::Integer = 1 << 20)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:tb_stream_begin, LIB), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), ctx(), capacity, h)
    rc == 0 || error("tb_stream_begin failed ($rc)")
    return BranchStream(h[], SlicedBranch[])
end

function stream_push!(st::BranchStream, branches::Vector{SlicedBranch})
    isempty(branches) && return st
    empty_net = TbNetwork(0, 0, C_NULL, C_NULL, 0, C_NULL, C_NULL, C_NULL, C_NULL, 0, 0, 0, 0, C_NULL, C_NULL)
    flats = [(nv(b.p.g) == 0 || isnothing(b.code)) ? nothing : FlatBranch(b) for b in branches]
    nets = TbNetwork[isnothing(f) ? empty_net : f.net for f in flats]
    GC.@preserve flats begin   # the library copies what it needs during the call
        rc = ccall((:tb_stream_push, LIB), Cint, (Ptr{Cvoid}, Ptr{TbNetwork}, Ptr{Float64}, Int64),
                   st.handle, nets, C_NULL, length(nets))
        rc == 0 || error("tb_stream_push failed ($rc): " *
                         unsafe_string(ccall((:tb_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx())))
    end
    append!(st.branches, branches)
    return st
end

function stream_finish!(st::BranchStream, element_type::Type)
    n = length(st.branches)
    vals = Vector{Float64}(undef, max(n, 1)); status = Vector{Int32}(undef, max(n, 1))
    got = Ref{Int64}(0); mx = Ref{Float64}(0)
    rc = ccall((:tb_stream_finish, LIB), Cint,
               (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int32}, Int64, Ref{Int64}, Ref{Float64}),
               st.handle, vals, status, max(n, 1), got, mx)
    st.handle = C_NULL
    rc == 0 || error("tb_stream_finish failed ($rc)")
    return element_type[(nv(b.p.g) == 0 || isnothing(b.code)) ? element_type(b.r) : element_type(vals[i]) + element_type(b.r)
                        for (i, b) in enumerate(st.branches)]
end

# branching_table(p, TensorNetworkSolver(), region) (src/branch.jl:79; default table solver src/types.jl:46): the table of a
# region with ALL optimal configurations per surviving row, in one C call (tb_branching_table: row optima, mis_compactify and
# the configuration rows on the device).  `vs` = the region's vertices, `ovs` = its open vertices (both as vertex ids of
# p.g); bit i-1 of a returned configuration is vertex vs[i], as in the reference's BranchingTable.  A maintainer hooks it in
# with `OptimalBranchingCore.branching_table(p::MISProblem, ::TensorNetworkSolver, vs) = BranchingTable(length(vs),
# TBCuda.branching_table_cuda(p.g, p.weights, vs, open_vertices(p.g, vs)))` (then prune_by_env as before, on the host).
struct FlatRegion   # keeps the arrays alive while the C call runs
    leaf_off::Vector{Int32}; leaf_labels::Vector{Int32}; weights::Any; boundary::Vector{Int32}; net::TbNetwork
end
function FlatRegion(g, weights, vs::Vector{Int}, ovs::Vector{Int})
    n = length(vs)
    n <= 32 || error("TBCuda: a region has at most 32 vertices")
    pos = Dict(v => Int32(i - 1) for (i, v) in enumerate(vs))
    leaf_off = Int32[0]; leaf_labels = Int32[]
    for v in vs                                   # vertex tensors
        push!(leaf_labels, pos[v]); push!(leaf_off, Int32(length(leaf_labels)))
    end
    for (i, u) in enumerate(vs), v in vs[i+1:end]  # edge tensors of the induced subgraph
        if TensorBranching.Graphs.has_edge(g, u, v)
            push!(leaf_labels, pos[u], pos[v]); push!(leaf_off, Int32(length(leaf_labels)))
        end
    end
    unit = weights isa TensorBranching.UnitWeight
    wv = unit ? nothing : collect(weights[vs])
    net = TbNetwork(n, length(leaf_off) - 1, pointer(leaf_off), pointer(leaf_labels), 0, C_NULL, C_NULL, C_NULL,
                    unit ? C_NULL : Ptr{Cvoid}(pointer(wv)), unit ? Int32(0) : weight_code(eltype(wv)), Int32(0), UInt32(0),
                    Int32(0), C_NULL, C_NULL)
    return FlatRegion(leaf_off, leaf_labels, wv, Int32[pos[v] for v in ovs], net)
end

"the tables of several regions [(vs, ovs), ...] of one graph in the same launches (tb_branching_tables); one row list per region"
function branching_tables_cuda(g, weights, regions::Vector{Tuple{Vector{Int}, Vector{Int}}})
    flats = [FlatRegion(g, weights, vs, ovs) for (vs, ovs) in regions]
    nets = [f.net for f in flats]
    boff = Int32[0]; blab = Int32[]
    for f in flats
        append!(blab, f.boundary); push!(boff, Int32(length(blab)))
    end
    base = cumsum([0; [1 << length(f.boundary) for f in flats]])
    nrows = base[end]
    keep = Vector{UInt8}(undef, nrows); sizes = Vector{Float64}(undef, nrows)
    row_off = Vector{Int64}(undef, nrows + 1); total = Ref{Int64}(0)
    cfgs = Vector{UInt32}(undef, max(4096, 64 * length(flats)))
    GC.@preserve flats begin
        for attempt in 1:2
            rc = ccall((:tb_branching_tables, LIB), Cint,
                       (Ptr{Cvoid}, Ptr{TbNetwork}, Ptr{Int32}, Ptr{Int32}, Int64, Ptr{UInt8}, Ptr{Float64}, Ptr{Int64}, Ptr{UInt32}, Int64, Ref{Int64}),
                       ctx(), nets, boff, blab, length(flats), keep, sizes, row_off, cfgs, length(cfgs), total)
            rc == 0 && break
            (attempt == 1 && total[] > length(cfgs)) || error("tb_branching_tables failed ($rc): " *
                unsafe_string(ccall((:tb_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx())))
            resize!(cfgs, total[])
        end
    end
    return [[cfgs[row_off[a]+1:row_off[a+1]] for a in base[i]+1:base[i+1] if keep[a] != 0] for i in eachindex(flats)]
end

branching_table_cuda(g, weights, vs::Vector{Int}, ovs::Vector{Int}) = branching_tables_cuda(g, weights, [(vs, ovs)])[1]

# The usecuda=true switch position.  The reference defines
#     contract_slices(::Vector{SlicedBranch}, ::Type, ::Bool)   and   solve_slice(::SlicedBranch, ::Type, ::Bool)
# (src/dynamic_ob.jl:30,36).  Re-defining exactly those signatures from here would OVERWRITE them (and `invoke` would then
# recurse), so these methods are strictly more specific in the element type -- `Type{T} where T<:AbstractFloat` wins the
# dispatch for Float32 / Float64 -- and the usecuda=false branch reaches the reference's own method through `invoke` with
# the original, less specific signature.  (A maintainer would rather add the two-line hook shown in INTEGRATION.md.)
# The device computes exactly in integers (unit / integer-valued weights, any float container), in Tropical{Float32}, or
# -- element_type Float64 with real weights -- in Tropical{Float64} (value_type_for).  Other float types stay with the reference.
integer_valued(w) = w isa TensorBranching.UnitWeight || all(isinteger, w)
on_device(branch::SlicedBranch, ::Type{T}) where {T} = T === Float32 || T === Float64 || integer_valued(branch.p.weights)

function TensorBranching.contract_slices(branches::Vector{SlicedBranch}, element_type::Type{T}, usecuda::Bool) where {T <: AbstractFloat}
    usecuda && all(b -> on_device(b, T), branches) && return contract_slices_cuda(branches, element_type)
    return invoke(TensorBranching.contract_slices, Tuple{Vector{SlicedBranch}, Type, Bool}, branches, element_type, false)
end
function TensorBranching.solve_slice(branch::SlicedBranch, element_type::Type{T}, usecuda::Bool) where {T <: AbstractFloat}
    (usecuda && on_device(branch, T)) || return invoke(TensorBranching.solve_slice, Tuple{SlicedBranch, Type, Bool}, branch, element_type, false)
    return contract_slices_cuda(SlicedBranch[TensorBranching.SlicedBranch(branch.p, branch.code, zero(branch.r))], element_type)[1]
end

end # module
