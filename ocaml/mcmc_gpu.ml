(* mcmc_gpu.ml -- OCaml side of the binding (see mcmc_gpu.mli).
   NOT COMPILED in this repository's image (no OCaml toolchain). *)
open Bigarray

type ctx
type tree

type logfn = { kind : int; dim : int; scale : float; params : (float, float64_elt, c_layout) Array1.t }
type proposal = { pkind : int; pdim : int; pparams : (float, float64_elt, c_layout) Array1.t }

external ctx_create : int -> int64 -> ctx = "mcmcgpu_ctx_create"
external set_seed : ctx -> int64 -> unit = "mcmcgpu_set_seed"
external reset_counters : ctx -> unit = "mcmcgpu_reset_counters"
external get_counters : ctx -> int * int = "mcmcgpu_get_counters"
external mcmc_array_raw :
  ctx -> logfn -> logfn -> proposal -> int -> int -> int -> int -> int ->
  (float, float64_elt, c_layout) Array2.t -> (float, float64_elt, c_layout) Array3.t -> unit
  = "mcmcgpu_mcmc_array_bytecode" "mcmcgpu_mcmc_array_native"
external interp_make_raw :
  ctx -> (float, float64_elt, c_layout) Array2.t -> (float, float64_elt, c_layout) Array1.t ->
  (float, float64_elt, c_layout) Array1.t -> tree = "mcmcgpu_interp_make"
external interp_jump_prob_raw :
  ctx -> tree -> int -> (float, float64_elt, c_layout) Array2.t -> (float, float64_elt, c_layout) Array1.t -> unit
  = "mcmcgpu_interp_jump_prob"
external interp_draw_raw : ctx -> tree -> int -> (float, float64_elt, c_layout) Array2.t -> unit = "mcmcgpu_interp_draw"
external stats_draw_raw : ctx -> int -> float -> float -> (float, float64_elt, c_layout) Array1.t -> unit = "mcmcgpu_stats_draw"
external posterior_indices_raw :
  ctx -> (float, float64_elt, c_layout) Array1.t -> (int64, int64_elt, c_layout) Array1.t -> unit = "mcmcgpu_nested_posterior_indices"
external harmonic_raw : ctx -> (float, float64_elt, c_layout) Array1.t -> float = "mcmcgpu_evidence_harmonic_mean"
external lebesgue_raw :
  ctx -> int -> float -> (float, float64_elt, c_layout) Array2.t -> (float, float64_elt, c_layout) Array1.t ->
  (float, float64_elt, c_layout) Array1.t -> float
  = "mcmcgpu_evidence_lebesgue_bytecode" "mcmcgpu_evidence_lebesgue_native"
external direct_raw :
  ctx -> int -> (float, float64_elt, c_layout) Array2.t -> (float, float64_elt, c_layout) Array1.t ->
  (float, float64_elt, c_layout) Array1.t -> float = "mcmcgpu_evidence_direct"

let create ?(device = 0) ?(seed = 0L) () = ctx_create device seed

let ba1 a = Array1.of_array float64 c_layout a
let ba2 rows = Array2.of_array float64 c_layout rows

(* kind ids of include/mcmc_gpu.h *)
let flat dim = { kind = 0; dim; scale = 1.0; params = ba1 [||] }
let box_prior ?(value = 0.0) lo hi =
  { kind = 2; dim = Array.length lo; scale = 1.0; params = ba1 (Array.concat [lo; hi; [|value|]]) }
let gaussian mu sigma = { kind = 4; dim = Array.length mu; scale = 1.0; params = ba1 (Array.append mu sigma) }
let gaussian_data data = { kind = 6; dim = 2; scale = 1.0; params = ba1 data }
let cauchy_data data = { kind = 7; dim = 2; scale = 1.0; params = ba1 data }
let box_proposal h = { pkind = 0; pdim = Array.length h; pparams = ba1 h }
let uniform_wrapping lo hi dx = { pkind = 1; pdim = Array.length lo; pparams = ba1 (Array.concat [lo; hi; dx]) }

let mcmc_array ctx ?(nbin = 0) ?(nskip = 1) ?(nchains = 1) n like prior prop start =
  let d = like.dim in
  let out = Array3.create float64 c_layout nchains n (d + 2) in
  mcmc_array_raw ctx like prior prop nbin nskip n nchains 0 (ba2 [|start|]) out;
  Array.init nchains (fun c ->
      Array.init n (fun s ->
          { Mcmc.value = Array.init d (fun i -> out.{c, s, i});
            like_prior = { Mcmc.log_likelihood = out.{c, s, d}; log_prior = out.{c, s, d + 1} } }))

module Interp = struct
  type interp_pdf = { tree : tree; dim : int }
  let make ctx pts low high = { tree = interp_make_raw ctx (ba2 pts) (ba1 low) (ba1 high); dim = Array.length low }
  let draw_high_level ctx n ip =
    let out = Array2.create float64 c_layout 1 ip.dim in
    interp_draw_raw ctx ip.tree n out;
    Array.init ip.dim (fun i -> out.{0, i})
  let draw ctx ip = draw_high_level ctx 0 ip
  let jump_prob_batch ctx ?(n = 0) ip pts =
    let out = Array1.create float64 c_layout (Array.length pts) in
    interp_jump_prob_raw ctx ip.tree n (ba2 pts) out;
    Array.init (Array.length pts) (fun i -> out.{i})
  let jump_prob_high_level ctx n ip _ pt = (jump_prob_batch ctx ~n ip [|pt|]).(0)
  let jump_prob ctx ip src pt = jump_prob_high_level ctx 0 ip src pt
end

module Evidence = struct
  let columns samples =
    let pts = ba2 (Array.map (fun s -> s.Mcmc.value) samples) in
    let ll = ba1 (Array.map (fun s -> s.Mcmc.like_prior.Mcmc.log_likelihood) samples) in
    let lp = ba1 (Array.map (fun s -> s.Mcmc.like_prior.Mcmc.log_prior) samples) in
    (pts, ll, lp)
  let evidence_harmonic_mean ctx samples = let _, ll, _ = columns samples in harmonic_raw ctx ll
  let evidence_lebesgue ctx ?(n = 64) ?(eps = 0.1) samples =
    let pts, ll, lp = columns samples in lebesgue_raw ctx n eps pts ll lp
  let evidence_direct ctx ?(n = 64) samples =
    let pts, ll, lp = columns samples in direct_raw ctx n pts ll lp
end

(* Stats.draw_uniform / draw_gaussian / draw_cauchy (stats.ml:89-91,113-128): n draws from the context's stream *)
let draw kind ctx a b n =
  let out = Array1.create float64 c_layout n in
  stats_draw_raw ctx kind a b out;
  Array.init n (fun i -> out.{i})
let draw_uniform ctx a b n = draw 0 ctx a b n
let draw_gaussian ctx mu sigma n = draw 1 ctx mu sigma n
let draw_cauchy ctx x0 gamma n = draw 2 ctx x0 gamma n

(* Nested.posterior_samples n output (nested.ml:167-178): the sampled points *)
let posterior_samples ctx n (pts : float array array) (log_wts : float array) =
  let lw = Array1.of_array float64 c_layout log_wts in
  let idx = Array1.create int64 c_layout n in
  posterior_indices_raw ctx lw idx;
  Array.init n (fun i -> pts.(Int64.to_int idx.{i}))
