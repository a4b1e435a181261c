/* hopb200.h -- C ABI of libhopb200.so: the B200 (sm_100a) implementation of hop's trajectory hot path.
 *
 * Becksteinlab/hop is pure Python and has no FFI; the boundary this library replaces is the body
 * of four Python methods.  Each entry point below names the reference code it stands in for
 * (paths relative to the hop source tree).  INTEGRATION.md shows the ctypes binding a hop
 * maintainer would add at those places.
 *
 * Conventions
 *   - every pointer documented as "device" is a CUDA device pointer owned by the caller
 *     (e.g. torch.Tensor.data_ptr() or hopb_dev_alloc); "host" pointers are ordinary memory;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); calls only
 *     enqueue work unless documented as synchronising;
 *   - every function returns 0 on success or a negative HOPB_E_* code; the message for the last
 *     error on a handle is hopb_last_error(h).  No C++ exception crosses this boundary;
 *   - one handle per GPU and per host thread; handles own only scratch memory;
 *   - grids are C-ordered [nx][ny][nz] (z fastest), coordinates are float32 [F][N][3] exactly as
 *     MDAnalysis' `AtomGroup.positions` yields them frame by frame.
 */
#ifndef HOPB200_H
#define HOPB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HOPB_OK 0
#define HOPB_E_INVALID (-1)        /* bad argument */
#define HOPB_E_CUDA (-2)           /* CUDA runtime error (see hopb_last_error) */
#define HOPB_E_CAPACITY (-3)       /* event buffer too small; *n_events tells how many were produced */
#define HOPB_E_TOO_MANY_SITES (-4) /* more than 32767 sites: not representable in hop's int16 map */
#define HOPB_E_NOMEM (-5)

#define HOPB_MAX_SITES 32767
#define HOPB_SUMMARY_INTS 11 /* int32 per (chunk, atom) in a summary buffer [chunk][11][N] */
#define HOPB_STATE_INTS 4    /* int32 per atom in a state buffer [4][N]: s0, t0, t1, on */

typedef struct hopb_handle hopb_handle;
typedef struct hopb_grid hopb_grid;

/* One site-to-site transition; what TransportNetwork._analyze_hops appends to
 * graph_properties[(s0, s)] (hop/graph.py:236-247).  Times are in frames. */
typedef struct hopb_event {
  int32_t frame;    /* frame at which site `s` was entered */
  int32_t iatom;    /* index into the tracked atom group */
  int32_t s0;       /* site left (may be -1: the reference's "(-1, N)" quirk, graph.py:217-221) */
  int32_t s;        /* site entered */
  int32_t tau;      /* t1 - t0: residence on s0 */
  int32_t tbarrier; /* frame - t1: time spent in the interstitial */
} hopb_event;

/* ---- library / handle -------------------------------------------------------------------- */
int hopb_version(void);
/* number of CUDA kernels this library has launched in this process (all handles) */
uint64_t hopb_launch_count(void);
int hopb_create(int device, hopb_handle** out);
void hopb_destroy(hopb_handle* h);
const char* hopb_last_error(const hopb_handle* h);

/* Plain device-memory helpers so that a binding needs nothing but this library. */
int hopb_dev_alloc(hopb_handle* h, uint64_t bytes, void** out);
int hopb_dev_free(hopb_handle* h, void* p);
int hopb_memset(hopb_handle* h, void* stream, void* dev, int value, uint64_t bytes);
int hopb_h2d(hopb_handle* h, void* stream, void* dev, const void* host, uint64_t bytes);
int hopb_d2h(hopb_handle* h, void* stream, void* host, const void* dev, uint64_t bytes);
int hopb_stream_sync(hopb_handle* h, void* stream);

/* ---- grid geometry ----------------------------------------------------------------------- */
/* Bin edges as numpy.histogramdd returned them in DensityCollector.init_histogram
 * (hop/density.py:99-117): three host float64 arrays with n+1 entries each.
 * `decimals` (host int[3], may be NULL) are the rounding decimals of hop/trajectory.py:435
 * computed by the caller; NULL lets the library compute int(-log10(min(diff(edges)))) + 6. */
int hopb_grid_create(hopb_handle* h, const double* edges_x, int nx, const double* edges_y, int ny,
                     const double* edges_z, int nz, const int* decimals, hopb_grid** out);
void hopb_grid_destroy(hopb_grid* g);

/* ---- stage 1: density ---------------------------------------------------------------------- */
/* counts[cell] += number of points in the cell; replaces the per-frame
 * `numpy.histogramdd(coord, bins, range)` + `grid += h` of DensityCollector.collect
 * (hop/density.py:125-131).  xyz: device float32 [n_points][3] (any number of whole frames);
 * counts: device uint32 [nx*ny*nz]. */
int hopb_hist_accumulate(hopb_handle* h, void* stream, const hopb_grid* g, const float* xyz,
                         int64_t n_points, uint32_t* counts);

/* Same, and additionally writes for every point the brick index of its cell under the
 * hopping-trajectory binning rules of hop/trajectory.py:428-441 (see hopb_brick_map; 0xFFFFFFFF =
 * outlier; bit 30 is bookkeeping of the slabbed histogram and ignored by hopb_hoptraj) so that
 * hopb_hoptraj need not read the coordinates again.
 * cells: device uint32 [n_points] (NULL = off).
 * Grids larger than the handle's L2 budget (hopb_set_hist_l2_budget, 64 MiB by default) are reduced in x-slab passes
 * over the cached brick indices, for calls with at least one point per cell through a grid of 16-bit counters packed two per word (half the bytes,
 * hence fewer passes) that is added into `counts` once a sum check has shown that no 16-bit counter wrapped; if one
 * did, the same call falls back to reductions straight into `counts`.  The result is exact either way. */
int hopb_hist_accumulate_cells(hopb_handle* h, void* stream, const hopb_grid* g, const float* xyz,
                               int64_t n_points, uint32_t* counts, uint32_t* cells);

/* Several selections (species) of one trajectory in ONE pass over the coordinates: hop runs one DensityCreator
 * per selection and so reads the trajectory once per species (scripts/hop-generate-many-densities.py:98-125,
 * hop/density.py:271-304), each selection on a grid of its own.  xyz: device float32 [F][N][3] holding all
 * selected atoms; species: device uint8 [N], the selection (0 .. n_species-1) of every atom; grids / counts:
 * HOST arrays of n_species grid handles and device count arrays (each += as hopb_hist_accumulate); cells: device
 * uint32 [F][N] or NULL, the brick index of every atom-frame in ITS species' grid (input of
 * hopb_hoptraj_strided).  n_species <= 8, regular (linspace) edges only. */
int hopb_hist_accumulate_multi(hopb_handle* h, void* stream, int n_species, const hopb_grid* const* grids,
                               const uint8_t* species, const float* xyz, int64_t F, int64_t N,
                               uint32_t* const* counts, uint32_t* cells);

/* Count grids larger than this many bytes (default 64 MiB) are histogrammed in x-slabs when `cells`
 * is given: one pass over the coordinates, then one pass over the 4-byte cell indices per further
 * slab, so that every pass reduces into an L2-resident part of the grid. */
int hopb_set_hist_l2_budget(hopb_handle* h, int64_t bytes);

/* Bulk density: the selection '<solvent> not within <cutoff> of <solute>' that
 * DensityCollector builds with notwithin_coordinates_factory (hop/density.py:82-87, an MDAnalysis
 * kd-tree search per frame), fused with the histogram.  Per frame, every solvent atom is counted in
 * counts_solvent (may be NULL) and, when no solute atom lies within the cutoff (d^2 <= cutoff^2 in
 * float32, no periodic images), also in counts_bulk -- the 'solvent' and 'bulk' collectors of
 * DensityCreator(mode='all') from one read of the coordinates.
 *   solvent  device float32 [F][Nw][3]   solute  device float32 [F][Np][3] (Np <= ~12000)
 *   cells    device uint32 [F][Nw] (may be NULL): brick indices of the solvent atoms, see above
 *   within   device uint8  [F][Nw] (may be NULL): 1 where a solute atom is within the cutoff */
int hopb_hist_bulk(hopb_handle* h, void* stream, const hopb_grid* g, const float* solvent, const float* solute,
                   int64_t F, int64_t Nw, int64_t Np, double cutoff, uint32_t* counts_solvent, uint32_t* counts_bulk,
                   uint32_t* cells, uint8_t* within);

/* density = counts / n_frames / dx[i] / dy[j] / dz[k] * factor, in float64 and in exactly that
 * order: DensityCollector.finish (hop/density.py:133-137), Grid.make_density
 * (hop/sitemap.py:196-216), Grid.convert_density (hop/sitemap.py:236-264).  Pass factor = 1.0 for
 * Angstrom^-3 (the multiplication is skipped, as in the reference).  density: device double [G]. */
int hopb_density_from_counts(hopb_handle* h, void* stream, const hopb_grid* g, const uint32_t* counts,
                             int64_t n_frames, double factor, double* density);

/* The same with the grid limited to ctas_per_sm CTAs per SM (0 = no limit): for running the density
 * beside other kernels on another stream without crowding them out of the SMs. */
int hopb_density_from_counts_ex(hopb_handle* h, void* stream, const hopb_grid* g, const uint32_t* counts,
                                int64_t n_frames, double factor, double* density, int ctas_per_sm);

/* ---- stage 2: site map --------------------------------------------------------------------- */
/* Density.map_sites (hop/sitemap.py:480-521): mask = density >= threshold; connected components
 * over the 7 forward offsets {0,1}^3 \ 0 and their negatives, no periodic wrap (_make_graph /
 * _shell, sitemap.py:1496-1541); labels 1..S by volume descending, ties by larger minimum
 * C-order cell index first (_label_connected_graphs, sitemap.py:1543-1571 with the canonical
 * tie-break of SURVEY.md F4); minsite == 2 drops single-cell components.
 *   labels   device int32 [G]  out: 0 = interstitial, 1..S
 *   n_sites  host   int32      out: S
 *   volumes  device int32 [HOPB_MAX_SITES+1] out (may be NULL): cells per label, volumes[0] = 0
 * Synchronises the stream (S is needed on the host). */
int hopb_label_sites(hopb_handle* h, void* stream, int nx, int ny, int nz, const double* density,
                     double threshold, int minsite, int32_t* labels, int32_t* n_sites, int32_t* volumes);

/* Same without the host round trip: n_sites_dev is a DEVICE int32 that receives S.  A value above
 * HOPB_MAX_SITES means "too many sites" (labels are then incomplete); the caller checks it whenever
 * it next synchronises. */
int hopb_label_sites_async(hopb_handle* h, void* stream, int nx, int ny, int nz, const double* density,
                           double threshold, int minsite, int32_t* labels, int32_t* n_sites_dev,
                           int32_t* volumes);

/* The same labelling straight from the histogram counts: "density >= threshold" is decided as
 * "count >= cmin(bin-width class of the cell)", cmin found on the device by bisection over the very
 * function hopb_density_from_counts evaluates (the density is monotone in the count, and the widths of
 * a linspace grid take a handful of distinct values per axis), so the labels are identical to
 * hopb_density_from_counts + hopb_label_sites_async while the density need not exist yet.
 * factor: the unit factor of hopb_density_from_counts.  Fails with HOPB_E_INVALID when
 * hopb_grid_width_classes(g) == 0. */
int hopb_label_sites_counts_async(hopb_handle* h, void* stream, const hopb_grid* g, const uint32_t* counts,
                                  int64_t n_frames, double factor, double threshold, int minsite,
                                  int32_t* labels, int32_t* n_sites_dev, int32_t* volumes);

/* Optional periodic wrapping in front of every binning that uses this grid (stage 1 and stage 3): each
 * coordinate is replaced by its image in [origin, origin + length) -- x - floor((x - o) * (1 / l)) * l in float32,
 * every step rounded separately -- before the bin search.  OFF BY DEFAULT, and hop itself never wraps: it bins
 * AtomGroup.positions as they are and asks for trajectories that were centred and wrapped beforehand
 * (hop/density.py:21-28; no minimum image in the neighbour graph either, hop/sitemap.py:1521-1525), so a drop-in
 * run leaves this off and the results are those of the reference bit for bit.  origin, length: host double[3];
 * NULL for either switches wrapping off again. */
int hopb_grid_set_periodic_box(hopb_grid* g, const double* origin, const double* length);

/* Number of distinct (dx, dy, dz) bin-width triples of the grid, or 0 when an axis has more than 16
 * distinct widths (irregular edges): then only the density-based labelling is available. */
int hopb_grid_width_classes(const hopb_grid* g);

/* Density.site_insert_bulk (hop/sitemap.py:877-918) on label grids: the cells with
 * bulk_labels == bulklabel become site 1, every site of `labels` is shifted by +1 and overwrites
 * the bulk where they overlap (_draw_map_from_sites, sitemap.py:1574-1587).
 *   own      device int32 [G] in/out: hydration-site label of the cell after the shift, else 0
 *   bulk_labels device int32 [G] (may be NULL = site_insert_nobulk, sitemap.py:933-943)
 *   member   device uint8 [G] out: 1 where the cell belongs to the inserted bulk site
 *   map16    device int16 [G] out: the painted map (what Density.map holds) */
int hopb_insert_bulk(hopb_handle* h, void* stream, int64_t n_cells, int32_t* own,
                     const int32_t* bulk_labels, int32_t bulklabel, uint8_t* member, int16_t* map16);

/* int32 labels -> int16 map without a bulk site (plain map_sites result). */
int hopb_labels_to_map16(hopb_handle* h, void* stream, int64_t n_cells, const int32_t* labels, int16_t* map16);

/* Density._annotate_sites (hop/sitemap.py:1589-1612, 601-627): per site label the cell count,
 * mean and population std of the density over the site's cells, and the mean cell midpoint.
 * Membership: label own[c] (if > 0) and, where member[c] != 0, additionally label 1.
 *   n_labels  number of labels including interstitial (len(sites))
 *   out_count device int64 [n_labels]; out_mean, out_std device double [n_labels];
 *   out_center device double [n_labels][3] (sum of midpoints / count; 0 for empty sites) */
int hopb_site_properties(hopb_handle* h, void* stream, const hopb_grid* g, const double* density,
                         const int32_t* own, const uint8_t* member, int n_labels, int64_t* out_count,
                         double* out_mean, double* out_std, double* out_center);

/* The whole of stage 2 in one (cooperative) launch, for the common workflow in which the bulk density
 * is the same histogram thresholded lower (no solute): what hop does with
 *   density = DensityCollector.finish / Grid.make_density / convert_density  (hop/density.py:133-137,
 *             hop/sitemap.py:196-216, 236-264)
 *   density.map_sites(solvent_threshold); bulk.map_sites(bulk_threshold)     (hop/sitemap.py:480-521,
 *             1496-1587)
 *   density.site_insert_bulk(bulk)                                           (hop/sitemap.py:877-918)
 *   _annotate_sites                                                          (hop/sitemap.py:1589-1612)
 * plus the stage-3 view of the map (hopb_brick_map).  Same labels as hopb_label_sites_counts_async +
 * hopb_insert_bulk; the annotation agrees with hopb_site_properties to rounding (one-pass shifted
 * variance).  Needs 1 <= hopb_grid_width_classes(g) <= 512 and nz <= 3072; HOPB_E_INVALID otherwise
 * (callers then use the stand-alone entry points).
 *   counts   device uint32 [nx*ny*nz]; n_frames, factor as hopb_density_from_counts
 *   with_bulk 0: plain map_sites (labels 1.., member untouched, may be NULL); 1: bulk site = label 1,
 *            hydration sites 2..
 *   density  device double [G] out; own device int32 [G] out; member device uint8 [G] out;
 *   map16    device int16 [G] out; brick / coarse / fine / cell2 as hopb_brick_map (fine, cell2 may be NULL)
 *   n_sites_dev device int32 [2] out: hydration sites, components of the bulk mask (a value above
 *            HOPB_MAX_SITES means the int16 map cannot hold them; all labels are then 0)
 *   volumes  device int32 [HOPB_MAX_SITES+1] or NULL: cells per hydration site, rank order
 *   p_*      annotation as hopb_site_properties over n_labels_cap labels, or all NULL */
int hopb_site_map_fused(hopb_handle* h, void* stream, const hopb_grid* g, const uint32_t* counts,
                        int64_t n_frames, double factor, double solvent_threshold, double bulk_threshold,
                        int with_bulk, int minsite, double* density, int32_t* own, uint8_t* member,
                        int16_t* map16, int16_t* brick, uint32_t* coarse, uint8_t* fine, uint32_t* cell2,
                        int32_t* n_sites_dev, int32_t* volumes, int n_labels_cap, int64_t* p_count,
                        double* p_mean, double* p_std, double* p_center);

/* Cell lists (Density.sites) in CSR form, cells ascending within a site.
 *   cells   device int32 [capacity]; offsets device int64 [n_labels+1]; n_cells host out.
 * Returns HOPB_E_CAPACITY (with *n_cells set) if capacity is too small.  Synchronises. */
int hopb_site_cells(hopb_handle* h, void* stream, int64_t n_grid_cells, const int32_t* own,
                    const uint8_t* member, int n_labels, int32_t* cells, int64_t capacity,
                    int64_t* offsets, int64_t* n_cells);

/* ---- rate constants (hop.graph.HoppingGraph._rate) ------------------------------ */
/* For every edge e: the survival function S(t) of its waiting times on the mesh t = 0, 1, ..., m_e - 1
 * (hop/graph.py:2433-2518; m_e = len(numpy.arange(0, max(tau) + 2, 1.0))) and the least-squares fit of
 * a exp(-k1 t) + (1 - a) exp(-k2 t) (more than 10 waiting times, start [0.5, 0.1, 1e-4]) or exp(-k t) (start
 * [1.0]; also when the double exponential is degenerate: |0.5 - a| > 0.49 or a negative k) by MINPACK's lmdif
 * with scipy.optimize.leastsq's defaults -- HoppingGraph._rate + fit_func (hop/graph.py:1677-1748, 2354-2419).
 *   tau      device double [off_tau[n_edges]], the waiting times edge after edge (any order inside an edge)
 *   off_tau  device int64 [n_edges + 1]; off_x device int64 [n_edges + 1], prefix sums of the mesh lengths
 *   perturb  0 for the reference's start values; otherwise they are multiplied by 1 + perturb
 *   noise    0 for the fit itself; != 0 nudges every exp() by +-1 ulp pseudo-randomly (seeded by `noise`) -- a second
 *            opinion: lmdif differentiates by forward differences with h = 1.5e-8 |p|, so rounding in the model
 *            function reaches the result at the 1e-8 .. 1e-6 level, more along flat directions.  A fit whose
 *            result moves visibly between noise = 0 and noise != 0 is not determined to the 1e-6 of the parity
 *            bar by ANY implementation other than the reference's own library call
 *   S        device double [total_x] out
 *   out      device double [n_edges][8] out: kind (0: exp, 1: double exp), the parameters p0, p1, p2, the rate
 *            constant k, MINPACK's info, its nfev, the number of waiting times */
int hopb_rate_fit(hopb_handle* h, void* stream, const double* tau, const int64_t* off_tau, const int64_t* off_x,
                  int64_t n_edges, int64_t total_x, double perturb, unsigned noise, double* S, double* out);

/* ---- per-site observables (hop.siteanalysis) ------------------------------ */
/* Atoms per site and frame: occ[f][l] = number of atoms whose label in frame f is l (outliers, -1, are not
 * counted) -- the per-frame numpy.histogram of Occupancy._compute_occupancy (hop/siteanalysis.py:837-841),
 * from the x plane of the hoptraj (occupancy) or its y plane (orbit occupancy, :860-861).
 *   labels device float32 [F][N]; occ device int32 [F][n_labels] out */
int hopb_site_occupancy(hopb_handle* h, void* stream, const float* labels, int64_t F, int64_t N, int n_labels,
                        int32_t* occ);

/* Distance of every selected atom from the centre of the site it is on, histogrammed per site: what
 * Distance._compute_distance + SiteHistogramArray.add do per frame (hop/siteanalysis.py:753-776, 531-544),
 * including that a (site, bin) pair hit by several atoms of one frame counts once for that frame, that the
 * outlier label -1 reads the last site's centre and row (Python negative index), and that NaN distances
 * (interstitial: centre NaN) fall into the upper outlier bin.
 *   xyz        device float32 [n_frames_total][N][3], coordinates of the selected atoms
 *   labels     device float32 [n_frames_total][N_hop], hoptraj x (Distance) or y (Orbit)
 *   atom_map   device int32 [N] hoptraj column of every selected atom (idcd2ihop), or NULL for identity
 *   frames     frame_start, frame_start + frame_step, ... < frame_stop
 *   centers    device double [n_labels][3] (site_properties.center, NaN for the interstitial)
 *   site2indx  device int32 [n_labels]: row of every label in hist; n_rows - 1 for sites not collected
 *   edges      device double [n_bins + 1]; hist device uint64 [n_rows][n_bins + 2], += (column 0 / n_bins + 1:
 *              below / above the range) */
int hopb_site_distance_hist(hopb_handle* h, void* stream, const float* xyz, const float* labels,
                            const int32_t* atom_map, int64_t N, int64_t N_hop, int64_t frame_start,
                            int64_t frame_stop, int64_t frame_step, const double* centers,
                            const int32_t* site2indx, int n_labels, const double* edges, int n_bins, int n_rows,
                            uint64_t* hist);

/* ---- stage 3 + 4: hopping trajectory and transition scan ------------------------------------ */
/* HoppingTrajectory._coord2hop (hop/trajectory.py:403-468) fused with the per-atom state machine
 * of TransportNetwork._analyze_hops (hop/graph.py:212-254) over one slab of F frames.
 * The slab is cut into n_chunks chunks of chunk_len frames (the last may be short); every
 * (atom, chunk) runs independently and leaves a summary, hopb_hop_fixup stitches them.
 *   xyz       device float32 [F][N][3]; ignored (may be NULL) when `cells` is given
 *   cells     device uint32 [F][N] (may be NULL): brick index of every atom-frame as written by
 *             hopb_hist_accumulate_cells -- the coordinates are then not read again
 *   brick     device int16: the site map (Density.map) in brick order, from hopb_brick_map
 *   coarse    device uint32 (may be NULL): the 2-bit block map of hopb_brick_map; it lets atoms in
 *             uniformly bulk / interstitial 4x4x4 bricks skip the gather from the site map
 *   fine      device uint8 (may be NULL): the 1-bit block maps of hopb_brick_map over 2x2x2 sub-bricks.
 *             Used instead of `coarse` whenever one of them fits into shared memory (grids up to
 *             ~230^3 cells): far fewer atoms then need the gather.  fine_value selects the map:
 *             0 (sub-bricks uniformly interstitial), 1 (uniformly bulk) or -1 = decided on the device
 *             (1 when the bulk site covers at least 1/8 of the grid).
 *   cell2     device uint32 (may be NULL): the 2-bit cell map of hopb_brick_map, the gather target for
 *             grids whose block maps do not fit into shared memory (beyond ~370^3 cells)
 *   frame0    global frame number of the slab's first frame
 *   hop_x/y   device float32 [F][N] out (site / orbit site: the X and Y records of the hoptraj
 *             DCD, trajectory.py:456-463); either may be NULL
 *   summaries device int32 [n_chunks][11][N] out
 *   events    device hopb_event [cap]; n_events device uint64 (running count, NOT reset here) */
int hopb_hoptraj(hopb_handle* h, void* stream, const hopb_grid* g, const float* xyz, const uint32_t* cells,
                 int64_t F, int64_t N, int64_t frame0, const int16_t* brick, const uint32_t* coarse,
                 const uint8_t* fine, int fine_value, const uint32_t* cell2, float* hop_x, float* hop_y, int n_chunks,
                 int64_t chunk_len, int32_t* summaries, hopb_event* events, uint64_t* n_events, uint64_t cap);

/* hopb_hoptraj for atoms that are a column range of wider arrays: xyz / cells / hop_x / hop_y point at the first
 * atom's column and consecutive frames are `ld` atoms apart (ld >= N; ld == N is hopb_hoptraj).  Several selections
 * binned by one hopb_hist_accumulate_multi pass are mapped with one call per selection on the shared cell array;
 * summaries and events stay per selection ([n_chunks][11][N], atom numbers 0 .. N-1). */
int hopb_hoptraj_strided(hopb_handle* h, void* stream, const hopb_grid* g, const float* xyz, const uint32_t* cells,
                         int64_t F, int64_t N, int64_t ld, int64_t frame0, const int16_t* brick,
                         const uint32_t* coarse, const uint8_t* fine, int fine_value, const uint32_t* cell2,
                         float* hop_x, float* hop_y, int n_chunks, int64_t chunk_len, int32_t* summaries,
                         hopb_event* events, uint64_t* n_events, uint64_t cap);

/* hopb_hoptraj_strided with the two planes as int16 [F][ld] (both required): the labels _coord2hop writes are values
 * of hop's int16 site map or -1 (hop/sitemap.py:518-519, hop/trajectory.py:456-463), so the conversion is lossless
 * and the kernel writes 4 instead of 8 bytes per atom-frame; the float32 of the DCD record is restored where a file
 * is written.  The whole-trajectory pipeline uses this form. */
int hopb_hoptraj_i16(hopb_handle* h, void* stream, const hopb_grid* g, const float* xyz, const uint32_t* cells,
                     int64_t F, int64_t N, int64_t ld, int64_t frame0, const int16_t* brick, const uint32_t* coarse,
                     const uint8_t* fine, int fine_value, const uint32_t* cell2, int16_t* hop_x, int16_t* hop_y,
                     int n_chunks, int64_t chunk_len, int32_t* summaries, hopb_event* events, uint64_t* n_events,
                     uint64_t cap);

/* Site map (device int16 [nx][ny][nz], Density.map) -> brick order + block maps.  A brick is a
 * 4x4x4 block of cells stored as 64 consecutive int16 (one 128-byte line), bricks in C order over
 * (ceil(nx/4), ceil(ny/4), ceil(nz/4)); inside a brick the cells are ordered by 2x2x2 sub-brick:
 * cell (i,j,k) lives at brick*64 + (i1<<5 | j1<<4 | k1<<3 | i0<<2 | j0<<1 | k0) with i1 = (i>>1)&1,
 * i0 = i&1, ...
 *   brick   device int16  [n_bricks*64] out (cells outside the grid are 0)
 *   coarse  device uint32 [ceil(n_bricks/16)] out: 2 bits per brick, 0 / 1 = all interstitial / all
 *           bulk, 3 = mixed
 *   fine    device uint8  [2 * nb_pad + 16] out, nb_pad = n_bricks rounded up to a multiple of 4; may be
 *           NULL.  Two maps: bit s of byte v * nb_pad + b is set when all cells of sub-brick
 *           s = (i1<<2 | j1<<1 | k1) of brick b carry the label v (0 = interstitial, 1 = bulk); the
 *           uint32 at byte 2 * nb_pad counts the cells with label 1.
 *   cell2   device uint32 [n_bricks * 4] out, 16-byte aligned, may be NULL: 2 bits per cell in brick
 *           order (cell index c: word c >> 4, bits 2 * (c & 15)): 0 = interstitial, 1 = bulk, 3 = any
 *           other label (look it up in `brick`) */
int hopb_brick_map(hopb_handle* h, void* stream, int nx, int ny, int nz, const int16_t* map16, int16_t* brick,
                   uint32_t* coarse, uint8_t* fine, uint32_t* cell2);

/* Same scan driven by an existing hopping trajectory (TransportNetwork on a loaded hoptraj,
 * hop/graph.py:190-254): labels = hop_x [F][N] float32. */
int hopb_hopscan_labels(hopb_handle* h, void* stream, const float* hop_x, int64_t F, int64_t N,
                        int64_t frame0, int n_chunks, int64_t chunk_len, int32_t* summaries,
                        hopb_event* events, uint64_t* n_events, uint64_t cap);

/* Fold the chunk summaries of one slab into a single summary [11][N] (the message ranks exchange). */
int hopb_hop_combine(hopb_handle* h, void* stream, int64_t N, int n_chunks, int64_t chunk_len, int64_t F,
                     const int32_t* summaries, int32_t* slab_summary);

/* state [4][N] (s0,t0,t1,on) <- apply n_summaries summaries [n][11][N] in order, no events.
 * With init != 0 the state starts from the virgin state (start of the trajectory). */
int hopb_hop_advance(hopb_handle* h, void* stream, int64_t N, int n_summaries, const int32_t* summaries,
                     int init, int32_t* state);

/* Stitch one slab: given the state at slab entry, emit the events the chunks left open, patch the
 * inherited orbit prefix of every chunk in hop_y (may be NULL), and leave the state at slab exit in
 * `state` (in/out, [4][N]). */
int hopb_hop_fixup(hopb_handle* h, void* stream, int64_t N, int n_chunks, int64_t chunk_len, int64_t F,
                   const int32_t* summaries, int32_t* state, float* hop_y, hopb_event* events,
                   uint64_t* n_events, uint64_t cap);

/* hopb_hop_fixup with hop_y rows `ld` atoms apart (see hopb_hoptraj_strided). */
int hopb_hop_fixup_strided(hopb_handle* h, void* stream, int64_t N, int64_t ld, int n_chunks, int64_t chunk_len,
                           int64_t F, const int32_t* summaries, int32_t* state, float* hop_y, hopb_event* events,
                           uint64_t* n_events, uint64_t cap);

/* hopb_hop_fixup_strided for an int16 hop_y plane (see hopb_hoptraj_i16). */
int hopb_hop_fixup_i16(hopb_handle* h, void* stream, int64_t N, int64_t ld, int n_chunks, int64_t chunk_len,
                       int64_t F, const int32_t* summaries, int32_t* state, int16_t* hop_y, hopb_event* events,
                       uint64_t* n_events, uint64_t cap);

/* Sort events into the reference's emission order and count transitions.
 * events_sorted is ordered by (s0, s) and within an edge by (frame, iatom) -- the order of the
 * per-edge lists of hop/graph.py:236-247.  The COO transition-count matrix is
 * (edge_s0[e], edge_s[e]) -> edge_count[e], with edge_start[e] the offset of the edge's first
 * event.  All outputs device; edge arrays need capacity min(n, (n_labels+1)^2) entries.
 * n_edges: host out.  Synchronises. */
int hopb_events_finalize(hopb_handle* h, void* stream, const hopb_event* events, uint64_t n, int64_t N,
                         int64_t F_total, int n_labels, hopb_event* events_sorted, int32_t* edge_s0,
                         int32_t* edge_s, int64_t* edge_start, int64_t* edge_count, int64_t* n_edges);

/* Same, and additionally fills a dense transition-count matrix: count_matrix (device int32
 * [(n_labels+1)^2], zeroed by the caller, may be NULL) gets count_matrix[(s0+1)*(n_labels+1) + (s+1)]
 * = number of (s0 -> s) transitions.  The whole finalisation is one cooperative kernel launch. */
int hopb_events_finalize_ex(hopb_handle* h, void* stream, const hopb_event* events, uint64_t n, int64_t N,
                            int64_t F_total, int n_labels, hopb_event* events_sorted, int32_t* edge_s0,
                            int32_t* edge_s, int64_t* edge_start, int64_t* edge_count, int32_t* count_matrix,
                            int64_t* n_edges);

/* The same finalisation with its sizes read on the device, so that the host can queue it behind
 * stage 3 without waiting for the event count or the number of sites (no synchronisation at all):
 *   n_events_dev  device uint64: number of events produced (the running count hopb_hoptraj adds to)
 *   cap           capacity of `events`, `events_sorted` and the four edge arrays
 *   n_sites_dev   device int32: number of hydration sites (hopb_label_sites_async);
 *                 n_labels = *n_sites_dev + label_offset (2 with a bulk site, else 1)
 *   count_matrix  device int32 [matrix_stride^2] or NULL: zeroed and filled here with row stride
 *                 matrix_stride (a fixed size, so that it can be all-reduced across ranks before the host
 *                 knows n_labels)
 *   result_dev    device int64 [4] out: n_events, n_edges, n_labels, status (bit mask of HOPB_FIN_*) */
#define HOPB_FIN_OVERFLOW 1        /* more events than `cap`: nothing was sorted; redo stage 3 with a larger buffer */
#define HOPB_FIN_TOO_MANY_SITES 2  /* *n_sites_dev > HOPB_MAX_SITES */
#define HOPB_FIN_MATRIX_SKIPPED 4  /* n_labels + 1 > matrix_stride: count_matrix was not filled */
int hopb_events_finalize_dev(hopb_handle* h, void* stream, const hopb_event* events, const uint64_t* n_events_dev,
                             uint64_t cap, int64_t N, int64_t F_total, const int32_t* n_sites_dev, int label_offset,
                             hopb_event* events_sorted, int32_t* edge_s0, int32_t* edge_s, int64_t* edge_start,
                             int64_t* edge_count, int32_t* count_matrix, int matrix_stride, int64_t* result_dev);

/* A plane of the hopping trajectory as int16 (lossless: the labels are the values of hop's int16 site map,
 * hop/sitemap.py:518-519, or -1), for a half-size download; the float32 of the DCD record (hop/trajectory.py:456-463)
 * is restored on the host.  labels: device float32 [n], 16-byte aligned; out: device int16 [n], 8-byte aligned. */
int hopb_labels_to_int16(hopb_handle* h, void* stream, const float* labels, int64_t n, int16_t* out);

/* TransportNetwork.graph_alltransitions (hop/graph.py:107-135): bit (a+1)*(n_labels+1)+(b+1) of
 * `bitmap` is set when some atom has label a in frame t-1 and b in frame t.
 * bitmap: device uint32 [ceil((n_labels+1)^2 / 32)], OR-ed into (not cleared). */
int hopb_all_transitions(hopb_handle* h, void* stream, const float* hop_x, int64_t F, int64_t N,
                         int n_labels, uint32_t* bitmap);

/* The same on an int16 plane (see hopb_hoptraj_i16). */
int hopb_all_transitions_i16(hopb_handle* h, void* stream, const int16_t* hop_x, int64_t F, int64_t N,
                             int n_labels, uint32_t* bitmap);

/* ---- exchange steps of the frame-sharded path on NVSwitch multicast memory ------ */
/* The reference is serial; sharding its frame loops (hop/density.py:125-131, hop/graph.py:212-254) over GPUs needs
 * three small exchanges: the sum of the count grids, the gather of one slab summary per atom and rank, the sum of
 * the transition-count matrices.  `mc` arguments are MULTICAST addresses of buffers in symmetric memory (the same
 * allocation on every rank, e.g. torch.distributed._symmetric_memory), valid for multimem.* instructions only; the
 * caller brackets every call with a cross-rank barrier before (all copies complete) and after (all slices
 * delivered).
 *
 * In-place sum over the ranks of n_words 32-bit counters: rank r reduces slice r through the switch
 * (multimem.ld_reduce.add) and writes it to every copy (multimem.st). */
int hopb_nvls_allreduce_u32(hopb_handle* h, void* stream, uint32_t* mc, int64_t n_words, int rank, int world);

/* The same sum with the additions on the SM: peer_ptrs = HOST array of `world` device pointers, the ranks' copies as
 * mapped into this process (peer mappings; entry `rank` is the local copy); slice `rank` is read from every copy with
 * 16-byte loads and the sum written to all copies with one multicast store. */
int hopb_p2p_allreduce_u32(hopb_handle* h, void* stream, const void* const* peer_ptrs, uint32_t* mc, int64_t n_words,
                           int rank, int world);

/* src (device, n_bytes, a multiple of 16) -> slot `rank` of every rank's [world][n_bytes] buffer behind mc_out. */
int hopb_nvls_allgather(hopb_handle* h, void* stream, const void* src, void* mc_out, int64_t n_bytes, int rank,
                        int world);

/* ---- synthetic input (bench / tests) ------------------------------------------- */
typedef struct hopb_synth_params {
  uint64_t seed;
  int64_t n_atoms, n_frames;
  int32_t site_stride, n_sites;
  float box, cube_half, sigma_site, sigma_bulk, p_hop, p_rebind;
} hopb_synth_params;

/* Writes frames [frame_start, frame_start+n_out) of the trajectory defined by `p` and the site
 * centres (device float32 [n_sites][3]) into xyz (device float32 [n_out][N][3]).  Bit-identical to
 * hop_b200/synth.py. */
int hopb_synth_generate(hopb_handle* h, void* stream, const hopb_synth_params* p, const float* centres,
                        int64_t frame_start, int64_t n_out, float* xyz);

/* The same for a trajectory produced chunk by chunk: `walker` (device float32 [n_atoms][4], 16-byte aligned)
 * receives the state of every atom after the last frame written; with resume != 0 the call continues from the
 * walker of the chunk that ended at frame_start - 1 instead of walking there from frame 0.  Same coordinates,
 * bit for bit, as one call over the whole range. */
int hopb_synth_generate_chunk(hopb_handle* h, void* stream, const hopb_synth_params* p, const float* centres,
                              int64_t frame_start, int64_t n_out, float* xyz, float* walker, int resume);

#ifdef __cplusplus
}
#endif
#endif /* HOPB200_H */
