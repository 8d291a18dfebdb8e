## nls_large_cuda_sparse.R -- R front-end of the B200 path of gsl_nls_large() for models with a SPARSE Jacobian.
##
## The reference takes the sparsity from a closure: `jac` (or the "gradient" attribute of `fn`) returns a
## dgCMatrix / dgRMatrix / dgTMatrix (R/nls_large.R:397-404, README "Sparse Jacobian matrix"), and the C side rebuilds
## a gsl_spmatrix from it on every callback (src/nls_large.c:528-623).  A closure cannot run on a GPU, so here the
## structure is DATA: the model is a list of blocks, each a row formula in a few local parameters plus, per local
## parameter, where it lives in the parameter vector.
##
##   ## README Example 4 (Penalty function I): fn <- function(theta) c(sqrt(1e-5) * (theta - 1), sum(theta^2) - 0.25)
##   p <- 500
##   fit <- gsl_nls_large_cuda_sparse(
##     blocks = list(
##       nls_block(~ sqrt(1e-5) * (th - 1), params = list(th = seq_len(p))),                   # rows 1..p
##       nls_block(~ th^2, params = list(th = seq_len(p)), rows = rep(p + 1L, p))),             # summed into row p+1
##     y = c(rep(0, p), 0.25), start = 1:p, algorithm = "cgst", control = list(maxiter = 500))
##
##   ## a grouped model, three nonzeros per Jacobian row: y ~ A[g] * exp(-lam * x) + b[g]
##   nls_block(~ A * exp(-lam * x) + b, params = list(A = g, lam = 2L * G + 1L, b = G + g), data = list(x = x))
##
## A parameter binding is one index (the same parameter for every term) or an integer vector with one (1-based)
## index per term.  Rows default to one row per term, blocks stacked in order; `rows` assigns terms to rows
## explicitly and a row is the SUM of its terms minus y.  algorithm = "cgst" (Steihaug-Toint) is matrix-free and
## has no size limit; "lm" (the reference's default), "dogleg", "ddogleg" and "subspace2D" factor a dense J^T J,
## which the library assembles from the nonzeros for up to 100 parameters (the reference densifies J for the same
## purpose, src/nls_large.c:641-648).

nls_block <- function(formula, params, data = list(), rows = NULL) {
  rhs <- formula[[length(formula)]]
  consts <- setdiff(all.vars(rhs), c(names(params), names(data)))
  if (length(consts))          # sqrt(1e-5) above is a call, not a variable; free symbols must be bound
    stop("unbound symbols in block formula: ", paste(consts, collapse = ", "))
  ## constant sub-expressions are folded in R so that the device sees the same double
  fold <- function(e) {
    if (is.call(e)) {
      e[-1L] <- lapply(as.list(e)[-1L], fold)
      if (!length(all.vars(e))) return(eval(e, baseenv()))
    }
    e
  }
  rhs_txt <- paste(deparse(fold(rhs), width.cutoff = 500L, control = "digits17"), collapse = " ")
  stopifnot(is.list(params), length(params) >= 1L, !is.null(names(params)), length(params) <= 16L)
  lens <- c(vapply(params, length, 1L), vapply(data, length, 1L), if (!is.null(rows)) length(rows))
  nterms <- max(lens)
  if (any(lens != nterms & lens != 1L) || any(vapply(data, length, 1L) != nterms))
    stop("variable lengths differ")
  scalar <- vapply(params, length, 1L) == 1L & nterms > 1L
  structure(list(
    rhs = rhs_txt, pnames = names(params),
    base = as.integer(ifelse(scalar, vapply(params, function(v) as.integer(v[1L]) - 1L, 1L), 0L)),
    index = lapply(seq_along(params), function(s) if (scalar[s]) NULL else as.integer(params[[s]]) - 1L),
    vnames = if (length(data)) names(data) else character(0L), cols = lapply(data, as.double),
    rows = if (is.null(rows)) NULL else as.integer(rows) - 1L, row0 = 0L, nterms = nterms), class = "nls_block")
}

gsl_nls_large_cuda_sparse <- function(blocks, y, start,
                                      algorithm = c("lm", "dogleg", "ddogleg", "subspace2D", "cgst"),
                                      weights = NULL, control = gsl_nls_control(),
                                      trace = FALSE, want_jtj = length(start) <= 1000L, device = 0L) {
  algorithm <- match.arg(algorithm)
  if (algorithm != "cgst" && length(start) > 100L)
    stop("more than 100 parameters: use algorithm = \"cgst\" (matrix-free); ", algorithm, " factors a dense J^T J")
  if (inherits(blocks, "nls_block")) blocks <- list(blocks)
  stopifnot(is.list(blocks), all(vapply(blocks, inherits, NA, "nls_block")))
  if (!is.numeric(start) || !length(start) || any(!is.finite(start)))
    stop("'start' must be a numeric vector of finite starting values")
  if (length(y) < length(start))
    stop("negative residual degrees of freedom, cannot fit a model with less observations than parameters")
  if (!is.null(weights) && (length(weights) != length(y) || any(!(weights > 0))))
    stop("'weights' should be a positive numeric vector of the same length as 'y'")
  ## stack blocks without explicit rows one after the other (R/nls_large.R: rows of fn in order)
  row0 <- 0L
  for (b in seq_along(blocks)) {
    blocks[[b]]$row0 <- row0
    if (is.null(blocks[[b]]$rows)) row0 <- row0 + blocks[[b]]$nterms
    if (any(unlist(blocks[[b]]$index) >= length(start)) || any(blocks[[b]]$base >= length(start)))
      stop("parameter index out of range in block ", b)
  }
  ctl <- .cuda_pack_control(control, algorithm, trace)
  cFit <- .Call(C_nls_large_cuda_sparse, blocks, as.double(y), as.double(start),
                if (is.null(weights)) NULL else as.double(weights), ctl$int, ctl$dbl,
                c(as.integer(isTRUE(want_jtj)), 1L), as.integer(device))
  names(cFit$par) <- names(start)
  convInfo <- list(isConv = as.logical(!cFit$conv), finIter = cFit$niter, finTol = cFit$ssrtol,
                   nEval = cFit$neval,
                   trsName = paste("multilarge", c(lm = "levenberg-marquardt", dogleg = "dogleg", ddogleg = "double-dogleg",
                                                   subspace2D = "2D-subspace", cgst = "steihaug-toint")[[algorithm]], sep = "/"),
                   stopCode = cFit$conv, stopMessage = cFit$status)
  m <- list(
    getPars = function() cFit$par, resid = function() cFit$resid, deviance = function() cFit$ssr,
    gradient = function() stop("the n x p Jacobian of a sparse model is not materialised; use $jtj or $grad_vec"),
    Rmat = function() if (is.null(cFit$jtj)) stop("fit with want_jtj = TRUE") else chol(cFit$jtj),
    conv = function() cFit$ssrtol, incr = function() cFit$ssrtol)
  class(m) <- "nlsModel"
  out <- list(m = m, convInfo = convInfo, control = ctl$ctrl, call = match.call(), weights = weights,
              ssrtrace = cFit$ssrtrace, grad_vec = cFit$grad_vec, jtj = cFit$jtj,
              cg_iters = cFit$cg_iters, nnz = cFit$nnz)
  class(out) <- c("gsl_nls", "nls")
  out
}
