## make_r_fixtures.R - pins the R-level pieces of the hot path's surroundings to a real R session.
##
## R is absent from the build image, so oracle/oracle_stats.py, oracle/jmb_oracle_mll.c and oracle/jmb_oracle_svft.c
## restate R's own code (density(), ar.yw, Hmisc::wtd.quantile, optim / vmmin / optimhess, nearPD) - "parity unpinned"
## in DESIGN.md.  A reviewer with R (>= 4.0), JMbayes (0.8-86), Hmisc and jsonlite runs, from the repository root,
##
##     Rscript tests/golden/make_r_fixtures.R
##
## which feeds the seeded inputs of tests/golden/r_inputs/*.json (written by tests/golden/export_r_inputs.py) to JMbayes'
## OWN functions and writes tests/golden/r_fixtures/*.json.  tests/test_r_fixtures.py then compares the restatements and
## the CUDA path with these numbers instead of skipping with "parity unpinned".
##
##   stats.json          JMbayes:::modes, :::effectiveSize, :::stdErr, :::computeP, stand / wmean / wsd as defined inside
##                       mvJointModelBayes() (R/mvJointModelBayes.R:939-947), quantile, Hmisc::wtd.quantile(type = "i/(n+1)")
##   marglogLik2.json    JMbayes:::marglogLik2(thetas, Data, priors, temp = 1, fixed_tau_Bs_gammas = TRUE / FALSE)
##                       (R/margLogLik2.R; it calls the package's compiled logPosterior / gradient_logPosterior)
##   survfit_modes.json  the mode search of survfitJM.mvJMbayes (R/survfitJM.mvJMbayes.R:295-303): optim(BFGS, hessian = TRUE)
##                       on -JMbayes:::log_post_RE_svft with the central-difference gradient JMbayes:::cd(eps = 1e-3)
suppressPackageStartupMessages({ library(jsonlite); library(JMbayes) })
inp <- file.path("tests", "golden", "r_inputs"); out <- file.path("tests", "golden", "r_fixtures")
dir.create(out, showWarnings = FALSE, recursive = TRUE)
M <- function (x) if (is.list(x) && !is.null(x$dim)) matrix(unlist(x$data), x$dim[[1]], x$dim[[2]]) else unlist(x)
wr <- function (x, name) write_json(x, file.path(out, name), digits = NA, auto_unbox = FALSE, na = "null")

## ---- posterior summaries ------------------------------------------------------------------------------------------
s <- fromJSON(file.path(inp, "stats.json"), simplifyVector = FALSE)
X <- M(s$X); LogLiks <- M(s$LogLiks)
stand <- function (x) { n <- length(x); upp <- max(x, na.rm = TRUE) + log(n); w <- exp(x - upp); w / sum(w) }
weights <- stand(LogLiks)
wmean <- function (x, weights, na.rm = FALSE) sum(x * weights, na.rm = na.rm)
wsd <- function (x, weights) sqrt(drop(cov.wt(cbind(x[!is.na(x)]), wt = weights)$cov))
tryNA <- function (expr) { r <- try(expr, silent = TRUE); if (inherits(r, "try-error")) NA else r }
rows <- lapply(seq_len(nrow(X)), function (i) { x <- X[i, ]
    list(postMeans = mean(x), postwMeans = wmean(x, weights), postModes = tryNA(JMbayes:::modes(x)),
         EffectiveSize = tryNA(unname(JMbayes:::effectiveSize(x))), StDev = sd(x), wStDev = tryNA(wsd(x, weights)),
         StErr = tryNA(unname(JMbayes:::stdErr(x))), CIs = unname(quantile(x, c(0.025, 0.975))),
         wCIs = tryNA(unname(Hmisc::wtd.quantile(x, weights = weights, probs = c(0.025, 0.975), type = "i/(n+1)", na.rm = TRUE))),
         Pvalues = JMbayes:::computeP(x)) })
wr(list(weights = weights, series = rows, R = R.version.string, JMbayes = as.character(packageVersion("JMbayes"))), "stats.json")

## ---- marglogLik2 ----------------------------------------------------------------------------------------------------
m <- fromJSON(file.path(inp, "marglogLik2.json"), simplifyVector = FALSE)
Data <- lapply(m$Data, M); Data$idGK_fast <- as.integer(unlist(m$Data$idGK_fast))
priors <- lapply(m$priors, M); thetas <- lapply(m$thetas, M)
res <- lapply(c(TRUE, FALSE), function (fixed) {
    th <- thetas; if (!fixed) th$tau_Bs_gammas <- log(th$tau_Bs_gammas)
    r <- JMbayes:::marglogLik2(th, Data, priors, temp = 1.0, fixed_tau_Bs_gammas = fixed)
    list(fixed = fixed, value = as.numeric(r), inits = attr(r, "inits"), Covs = attr(r, "Covs")) })
wr(res, "marglogLik2.json")

## ---- survfitJM mode search ----------------------------------------------------------------------------------------
v <- fromJSON(file.path(inp, "survfit_modes.json"), simplifyVector = FALSE)
q <- v$q; K <- v$K
modes <- lapply(v$subjects, function (sb) {
    Data <- list(idL = lapply(sb$y, function (y) rep(1L, length(unlist(y)))), idL2 = lapply(sb$y, function (y) length(unlist(y))),
                 idGK = K - 1L, idTs = lapply(sb$ZZs, function (z) rep(1L, K)), Us = lapply(sb$Us, M), y = lapply(sb$y, M),
                 Xbetas = lapply(sb$Xbetas, M), Z = lapply(sb$Z, M), RE_inds = lapply(v$RE_inds, unlist),
                 RE_inds2 = lapply(v$RE_inds2, unlist), W1s = M(sb$W1s), W2s = M(sb$W2s), XXsbetas = lapply(sb$XXsbetas, M),
                 ZZs = lapply(sb$ZZs, M), col_inds = lapply(v$col_inds, unlist), row_inds_Us = seq_len(K),
                 Bs_gammas = M(v$Bs_gammas), gammas = M(v$gammas), alphas = M(v$alphas), fams = unlist(v$fams),
                 links = unlist(v$links), sigmas = lapply(v$sigmas, function (x) if (is.null(x)) NA_real_ else x),
                 invD = M(v$invD), Pw = M(sb$Pw), trans_Funs = unlist(v$trans_Funs))
    ff <- function (b, Data) -JMbayes:::log_post_RE_svft(b, Data = Data)
    gg <- function (b, Data) JMbayes:::cd(b, ff, Data = Data, eps = 1e-03)
    opt <- optim(rep(0, q), ff, gg, Data = Data, method = "BFGS", hessian = TRUE,
                 control = list(maxit = 200, parscale = rep(0.1, q)))
    list(par = opt$par, value = opt$value, counts = unname(opt$counts), convergence = opt$convergence, hessian = opt$hessian) })
wr(modes, "survfit_modes.json")
cat("fixtures written to", out, "\n")
